"""Mint tests/golden/byte_level.json by running the REFERENCE's convert_to_byte_level (build container only) on the
synthetic tokenizers of bytelevel_cases.py.  Stored per case: the converted tokenizer's JSON, n_added, encodings of a few
sample texts, and the special-token attributes."""
import json
import os
import sys
from unittest.mock import MagicMock

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
for n in ["jax", "jax.numpy", "jax.sharding", "flax", "flax.linen", "flax.serialization", "flax.traverse_util", "optax"]:
    sys.modules.setdefault(n, MagicMock())
sys.modules["flax.linen"].Module = type("Module", (), {})
sys.path.insert(0, "/root/reference")

import bytelevel_cases  # noqa: E402
from zett.tokenizer_converters import convert_to_byte_level as ref_convert  # noqa: E402


def describe(tok, n_added):
    return {
        "tokenizer_json": json.loads(tok._tokenizer.to_str()),
        "n_added": n_added,
        "tokens": tok.convert_ids_to_tokens(range(len(tok))),
        "encodings": {t: tok.encode(t) for t in bytelevel_cases.SAMPLE_TEXTS},
        "specials": {k: getattr(tok, k) for k in ("bos_token", "eos_token", "unk_token", "pad_token", "sep_token", "cls_token", "mask_token")},
    }


def main():
    out = {}
    for name, (build, kwargs) in bytelevel_cases.cases().items():
        tok, n_added = ref_convert(build(), **kwargs())
        out[name] = describe(tok, n_added)
        print(name, len(tok), n_added)
    json.dump(out, open(os.path.join(HERE, "byte_level.json"), "w"), ensure_ascii=False)


if __name__ == "__main__":
    main()
