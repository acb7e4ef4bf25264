"""Mint the golden fixtures in this directory by running the REFERENCE ITSELF (build container only).

    python tests/golden/make_golden.py

* ``hypernet_<case>.npz``      -- outputs of ``/root/reference/hf_hypernet`` ``ZettHypernet.__call__`` (fp32,
  eager attention) on seeded synthetic weights / surface forms (zett_b200.synthetic generators).
* ``surface_forms_<case>.npz`` -- outputs of ``/root/reference/zett/utils.py:get_surface_form_matrix`` on
  seeded synthetic Unigram / BPE hn-tokenizers and byte-level target vocabularies.

``/root/reference`` cannot travel to the GPU box, so the vectors are committed; this script is the
provenance record.  The reference needs two stubs to import offline:
``RobertaConfig.from_pretrained`` (hub fetch of roberta-base -> its constants) and the jax/flax/optax
modules that ``zett/utils.py`` imports at module top but the surface-form function never touches.
"""
import json
import os
import sys
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REFERENCE = "/root/reference"

import zett_synthetic as synthetic  # noqa: E402


def load_reference_hypernet():
    import torch  # noqa: F401
    from transformers import RobertaConfig

    RobertaConfig.from_pretrained = classmethod(lambda cls, *a, **k: RobertaConfig(
        vocab_size=50265, max_position_embeddings=514, type_vocab_size=1, layer_norm_eps=1e-5,
        pad_token_id=1, bos_token_id=0, eos_token_id=2))
    sys.path.insert(0, REFERENCE)
    from hf_hypernet.configuration_hypernet import ZettHypernetConfig as RefConfig
    from hf_hypernet.modeling_hypernet import ZettHypernet as RefHypernet
    return RefConfig, RefHypernet


def load_reference_utils():
    for n in ["jax", "jax.numpy", "jax.sharding", "flax", "flax.linen", "flax.serialization",
              "flax.traverse_util", "optax"]:
        sys.modules.setdefault(n, MagicMock())
    sys.modules["flax.linen"].Module = type("Module", (), {})
    sys.path.insert(0, REFERENCE)
    import zett.utils as ref_utils
    return ref_utils


HYPERNET_CASES = {
    # case: (synthetic config name, overrides, n_rows, lang_index)
    "tiny": ("tiny", {}, 64, None),
    "tiny_lang": ("tiny_lang", {}, 64, 3),
    "tiny_single_head": ("tiny", {"hn_single_head": True}, 48, None),
    "tiny_plain": ("tiny", {"hn_rescale_embeddings": False, "hn_predict_bias": False,
                            "separate_out_embeddings": False, "hn_n_extra_tokens": 0}, 48, None),
}


def run_reference_hypernet(cfg, weights, surface_forms, source_embeddings, lang_index):
    import torch
    RefConfig, RefHypernet = load_reference_hypernet()
    ref_cfg = RefConfig(**{k: v for k, v in cfg.to_dict().items()
                           if k.startswith("hn_") or k in ("n_embd", "n_langs", "pad_token_id", "original_vocab_size",
                                                           "separate_out_embeddings", "use_unigram_bias")})
    model = RefHypernet(ref_cfg).eval()
    model.model.config._attn_implementation = "eager"
    missing, unexpected = model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()}, strict=False)
    assert not unexpected, unexpected
    assert all("position_ids" in m or "token_type_ids" in m for m in missing), missing
    with torch.no_grad():
        out = model(torch.from_numpy(surface_forms), source_embeddings=torch.from_numpy(source_embeddings),
                    lang_index=None if lang_index is None else torch.tensor(lang_index))
    return [None if o is None else o.numpy() for o in out]


def mint_hypernet():
    for case, (name, overrides, n_rows, lang_index) in HYPERNET_CASES.items():
        cfg = synthetic.make_config(name, **overrides)
        weights = synthetic.make_weights(cfg, seed=11)
        src = synthetic.make_source_embeddings(cfg, seed=12)
        sf = synthetic.make_random_surface_forms(cfg, n_rows, seed=13)
        pred_in, pred_out, bias = run_reference_hypernet(cfg, weights, sf, src, lang_index)
        checksum = float(sum(v.astype(np.float64).sum() for v in weights.values()) + src.astype(np.float64).sum())
        arrays = dict(surface_forms=sf, pred_in=pred_in, pred_bias=bias,
                      meta=np.array(json.dumps(dict(config=name, overrides=overrides, weight_seed=11, source_seed=12,
                                                    lang_index=lang_index, input_checksum=checksum))))
        if pred_out is not None:
            arrays["pred_out"] = pred_out
        np.savez_compressed(os.path.join(HERE, f"hypernet_{case}.npz"), **arrays)
        print("hypernet", case, pred_in.shape, None if pred_out is None else pred_out.shape, bias.shape)


def _drop_chars(vocab, scores, chars):
    keep = [i for i, t in enumerate(vocab) if t not in chars]
    return [vocab[i] for i in keep], scores[keep]


def mint_surface_forms():
    from tokenizers import Tokenizer, models
    from transformers import PreTrainedTokenizerFast
    ref_utils = load_reference_utils()

    vocab, scores = synthetic.make_hn_vocab(2000, seed=21)
    # integer-valued scores on a slice of the vocabulary force Viterbi ties
    scores = scores.copy()
    scores[300:900] = np.round(scores[300:900])
    # remove a few alphabet chars so unknown characters (and unk fusing) occur
    dropped = [synthetic.BYTES_TO_CHARS[b] for b in (0x71, 0x7A, 0xC3, 0x00)]  # 'q', 'z', 'Ã', 'Ā'
    vocab_u, scores_u = _drop_chars(vocab, scores, dropped)

    targets = synthetic.make_target_tokens(3000, seed=22, specials=("</s>", "<pad>"))
    rng = np.random.default_rng(23)
    alphabet = [synthetic.BYTES_TO_CHARS[b] for b in range(256)]
    for _ in range(200):  # arbitrary byte strings incl. 2-byte chars and long tokens (truncation)
        n = int(rng.integers(1, 24))
        targets.append("".join(alphabet[int(b)] for b in rng.integers(0, 256, size=n)))
    for i in range(0, 600, 3):  # concatenations of pieces -> multi-piece segmentations
        targets.append(vocab[300 + i] + vocab[301 + i] + vocab[302 + i])
    targets.append("<unk>")

    cases = {}
    uni = models.Unigram([(t, float(s)) for t, s in zip(vocab_u, scores_u)], unk_id=3, byte_fallback=False)
    cases["unigram"] = (uni, dict(type="unigram", vocab=vocab_u, scores=scores_u.tolist(), unk_id=3,
                                  byte_fallback=False), "<pad>")
    # BPE merges need every merge operand in the vocabulary: only drop chars no piece contains
    dropped_b = [synthetic.BYTES_TO_CHARS[b] for b in (0x21, 0x41, 0xC3, 0x00)]  # '!', 'A', 'Ã', 'Ā'
    vocab_b, _ = _drop_chars(vocab, scores, dropped_b)
    bvocab, merges = synthetic.make_bpe_merges(vocab_b)
    bpe = models.BPE(vocab={t: i for i, t in enumerate(bvocab)}, merges=merges, unk_token="<unk>")
    cases["bpe"] = (bpe, dict(type="bpe", vocab=bvocab, merges=[list(m) for m in merges], unk_token="<unk>",
                              fuse_unk=False, byte_fallback=False, ignore_merges=False), "</s>")
    bpe_f = models.BPE(vocab={t: i for i, t in enumerate(bvocab)}, merges=merges, unk_token="<unk>", fuse_unk=True,
                       ignore_merges=True)
    cases["bpe_fuse_ignore"] = (bpe_f, dict(type="bpe", vocab=bvocab, merges=[list(m) for m in merges],
                                            unk_token="<unk>", fuse_unk=True, byte_fallback=False,
                                            ignore_merges=True), "</s>")

    for case, (model, spec, pad_token) in cases.items():
        hn = PreTrainedTokenizerFast(tokenizer_object=Tokenizer(model), bos_token="<s>", pad_token=pad_token,
                                     eos_token="</s>", unk_token="<unk>")
        matrix, n_trunc = ref_utils.get_surface_form_matrix(list(targets), maxlen=7, tokenizer_to_use=hn, padding=5)
        spec["pad_token_id"] = int(hn.pad_token_id)
        spec["special_tokens"] = {t: int(hn.convert_tokens_to_ids(t)) for t in hn.all_special_tokens}
        np.savez_compressed(os.path.join(HERE, f"surface_forms_{case}.npz"), matrix=matrix,
                            n_truncated=np.int64(n_trunc), tokens=np.array(json.dumps(targets)),
                            spec=np.array(json.dumps(spec)), padding=np.int64(5), maxlen=np.int64(7))
        print("surface forms", case, matrix.shape, "truncated", n_trunc,
              "hist", synthetic.length_histogram(matrix[:-5], hn.pad_token_id))


if __name__ == "__main__":
    mint_surface_forms()
    mint_hypernet()
