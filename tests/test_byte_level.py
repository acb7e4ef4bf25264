"""zett_b200.byte_level.convert_to_byte_level against tokenizers converted by the REFERENCE itself
(tests/golden/byte_level.json, minted by tests/golden/make_golden_bytelevel.py on the inputs of bytelevel_cases.py)."""
import json
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import bytelevel_cases  # noqa: E402

from zett_b200.byte_level import convert_to_byte_level  # noqa: E402

GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "byte_level.json")))
CASES = bytelevel_cases.cases()


@pytest.mark.parametrize("case", sorted(CASES))
def test_matches_reference(case):
    build, kwargs = CASES[case]
    tok, n_added = convert_to_byte_level(build(), **kwargs())
    want = GOLDEN[case]
    assert n_added == want["n_added"]
    # the id -> byte-level token list is what get_surface_form_matrix consumes: identical
    assert tok.convert_ids_to_tokens(range(len(tok))) == want["tokens"]
    got = json.loads(tok._tokenizer.to_str())
    ref = want["tokenizer_json"]
    for key in ("normalizer", "pre_tokenizer", "post_processor", "decoder"):
        assert got.get(key) == ref.get(key), key
    gm, rm = got["model"], ref["model"]
    assert gm["type"] == rm["type"]
    if gm["type"] == "BPE":
        assert gm["vocab"] == rm["vocab"]
        norm = lambda ms: [tuple(m.split(" ")) if isinstance(m, str) else tuple(m) for m in ms]  # noqa: E731
        g, r = norm(gm["merges"]), norm(rm["merges"])
        assert sorted(g) == sorted(r) and len(set(g)) == len(set(r))
        # the original merges keep their relative order; only the repair merges (whose order the reference takes from
        # a set walk) may be permuted among themselves
        ws = lambda m: sum(c in "ĠĊĉ" for c in m[0] + m[1])  # noqa: E731
        inputs = set(norm(json.loads(build()._tokenizer.to_str())["model"]["merges"]))
        stable = [m for m in r if m in inputs and ws(m) <= 1 and r.count(m) == 1]
        assert [m for m in g if m in set(stable)] == stable
    else:
        assert gm["vocab"] == (rm["vocab"] if gm["type"] != "Unigram" else [list(v) for v in rm["vocab"]]) or gm["vocab"] == rm["vocab"]
    for k, v in want["specials"].items():
        assert getattr(tok, k) == v, k
    for text, ids in want["encodings"].items():
        assert tok.encode(text) == ids, text


def test_surface_forms_from_converted_tokenizer():
    """End of the host pipeline: converted target tokenizer -> get_surface_form_matrix (scripts/transfer.py:198-206)."""
    import numpy as np
    from oracle import retok_oracle as ro
    import zett_synthetic as synthetic
    from zett_b200.surface_forms import get_surface_form_matrix
    hn = synthetic.make_hn_tokenizer("unigram", 1200, seed=3)
    target, _ = convert_to_byte_level(bytelevel_cases.bpe_metaspace())
    got, nt = get_surface_form_matrix(target, 7, hn)
    want, nt2 = ro.surface_form_matrix_hf(target.convert_ids_to_tokens(range(len(target))), 7, hn)
    np.testing.assert_array_equal(got, want)
    assert nt == nt2 and got.shape[0] == len(target)
