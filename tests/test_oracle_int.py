"""The pure-Python retokenizer oracle (oracle/retok_oracle.py) against (a) goldens minted by running the
reference's get_surface_form_matrix and (b) the installed HF ``tokenizers`` wheel on randomised cases."""
import json
import os

import numpy as np
import pytest

from oracle import retok_oracle as ro
import zett_synthetic as synthetic

INT_CASES = ["unigram", "bpe", "bpe_fuse_ignore"]


def load_int_case(golden_dir, case):
    g = np.load(os.path.join(golden_dir, f"surface_forms_{case}.npz"))
    return g, json.loads(str(g["spec"])), json.loads(str(g["tokens"]))


def oracle_model(spec):
    if spec["type"] == "unigram":
        return ro.UnigramOracle(list(zip(spec["vocab"], spec["scores"])), spec["unk_id"], spec["byte_fallback"])
    return ro.BPEOracle({t: i for i, t in enumerate(spec["vocab"])}, [tuple(m) for m in spec["merges"]],
                        unk_token=spec["unk_token"], fuse_unk=spec["fuse_unk"], byte_fallback=spec["byte_fallback"],
                        ignore_merges=spec["ignore_merges"])


def test_byte_table_matches_synthetic():
    assert ro.CHARS_TO_BYTES == synthetic.CHARS_TO_BYTES
    assert len(ro.CHARS_TO_BYTES) == 256
    assert ro.BYTES_TO_CHARS[32] == "Ġ" and ro.BYTES_TO_CHARS[0] == "Ā" and ro.BYTES_TO_CHARS[173] == "Ń"


@pytest.mark.parametrize("case", INT_CASES)
def test_oracle_matches_reference_goldens(golden_dir, case):
    g, spec, tokens = load_int_case(golden_dir, case)
    matrix, n_trunc = ro.surface_form_matrix(tokens, int(g["maxlen"]), oracle_model(spec), spec["pad_token_id"],
                                             spec["special_tokens"], padding=int(g["padding"]))
    assert matrix.dtype == np.int32 and matrix.shape == g["matrix"].shape
    assert n_trunc == int(g["n_truncated"])
    np.testing.assert_array_equal(matrix, g["matrix"])


def test_key_error_outside_alphabet(golden_dir):
    g, spec, tokens = load_int_case(golden_dir, "unigram")
    with pytest.raises(KeyError):
        ro.surface_form_matrix(["ab c"], 7, oracle_model(spec), 1, {})


def _random_cases(rng, alphabet, n, maxlen=12):
    return ["".join(alphabet[int(i)] for i in rng.integers(0, len(alphabet), size=int(rng.integers(0, maxlen))))
            for _ in range(n)]


def test_unigram_vs_hf_tokenizers_randomised():
    from tokenizers import models
    rng = np.random.default_rng(5)
    alphabet = ["a", "b", "c", "Ġ", "é"]
    for trial in range(30):
        pieces = {"<unk>"}
        for _ in range(int(rng.integers(5, 40))):
            pieces.add("".join(alphabet[int(i)] for i in rng.integers(0, 4, size=int(rng.integers(1, 5)))))
        pieces = sorted(pieces)
        # integer scores force ties; duplicates exercise the last-wins id map
        vocab = [(p, float(-rng.integers(1, 6))) for p in pieces]
        if trial % 3 == 0:
            vocab.append((vocab[2][0], -1.0))
        unk_id = pieces.index("<unk>")
        hf = models.Unigram(vocab, unk_id=unk_id, byte_fallback=False)
        mine = ro.UnigramOracle(vocab, unk_id)
        for s in _random_cases(rng, alphabet, 300):
            assert [t.id for t in hf.tokenize(s)] == mine.tokenize(s), (trial, s)


def test_unigram_byte_fallback_vs_hf():
    from tokenizers import models
    vocab = [("<unk>", 0.0)] + [("<0x%02X>" % b, -5.0) for b in range(256)] + [("ab", -1.0), ("a", -2.0), ("b", -2.5)]
    hf = models.Unigram(vocab, unk_id=0, byte_fallback=True)
    mine = ro.UnigramOracle(vocab, 0, byte_fallback=True)
    for s in ["abc", "éab", "xyzab", "", "abĠĠa"]:
        assert [t.id for t in hf.tokenize(s)] == mine.tokenize(s), s


def test_bpe_vs_hf_tokenizers_randomised():
    from tokenizers import models
    rng = np.random.default_rng(6)
    alphabet = ["a", "b", "c", "Ġ", "é"]
    for trial in range(30):
        vocab = {"<unk>": 0}
        for ch in alphabet[:4]:
            vocab[ch] = len(vocab)
        merges = []
        symbols = list(alphabet[:4])
        for _ in range(int(rng.integers(3, 30))):
            a, b = symbols[int(rng.integers(len(symbols)))], symbols[int(rng.integers(len(symbols)))]
            if (a, b) in merges or len(a + b) > 6:
                continue
            merges.append((a, b))
            if a + b not in vocab:
                vocab[a + b] = len(vocab)
                symbols.append(a + b)
        for fuse in (False, True):
            hf = models.BPE(vocab=vocab, merges=merges, unk_token="<unk>", fuse_unk=fuse)
            mine = ro.BPEOracle(vocab, merges, unk_token="<unk>", fuse_unk=fuse)
            for s in _random_cases(rng, alphabet, 200):
                assert [t.id for t in hf.tokenize(s)] == mine.tokenize(s), (trial, fuse, s)
        hf = models.BPE(vocab=vocab, merges=merges)  # no unk token: unknown chars are dropped
        mine = ro.BPEOracle(vocab, merges)
        for s in _random_cases(rng, alphabet, 100):
            assert [t.id for t in hf.tokenize(s)] == mine.tokenize(s), (trial, s)


def test_bpe_prefix_suffix_vs_hf():
    from tokenizers import models
    vocab = {"<unk>": 0, "a": 1, "##a": 2, "##b": 3, "b": 4, "##b</w>": 5, "ab": 6, "##a</w>": 7, "a</w>": 8,
             "b</w>": 9, "ab</w>": 10, "##ab": 11}
    merges = [("a", "##b"), ("a", "##b</w>"), ("##a", "##b")]
    hf = models.BPE(vocab=vocab, merges=merges, unk_token="<unk>", continuing_subword_prefix="##",
                    end_of_word_suffix="</w>")
    mine = ro.BPEOracle(vocab, merges, unk_token="<unk>", continuing_subword_prefix="##", end_of_word_suffix="</w>")
    for s in ["ab", "aab", "abab", "a", "b", "ba", "abc", "cab"]:
        assert [t.id for t in hf.tokenize(s)] == mine.tokenize(s), s
