"""Native retokenizer (C ABI, host code) against the reference-minted goldens, the Python oracle and the HF
``tokenizers`` wheel; bit-exact."""
import json
import os

import numpy as np
import pytest

from oracle import retok_oracle as ro
import zett_synthetic as synthetic
from zett_b200.surface_forms import NativeTokenizerModel, get_surface_form_matrix

INT_CASES = ["unigram", "bpe", "bpe_fuse_ignore"]


def native_model(spec):
    if spec["type"] == "unigram":
        return NativeTokenizerModel.unigram(list(zip(spec["vocab"], spec["scores"])), spec["unk_id"], spec["byte_fallback"])
    return NativeTokenizerModel.bpe({t: i for i, t in enumerate(spec["vocab"])}, [tuple(m) for m in spec["merges"]],
                                    unk_token=spec["unk_token"], fuse_unk=spec["fuse_unk"],
                                    byte_fallback=spec["byte_fallback"], ignore_merges=spec["ignore_merges"])


@pytest.mark.parametrize("case", INT_CASES)
@pytest.mark.parametrize("threads", [1, 5])
def test_native_matches_reference_goldens(golden_dir, case, threads):
    g = np.load(os.path.join(golden_dir, f"surface_forms_{case}.npz"))
    spec, tokens = json.loads(str(g["spec"])), json.loads(str(g["tokens"]))
    sp = np.array([spec["special_tokens"].get(t, -1) for t in tokens], dtype=np.int32)
    out, n_trunc = native_model(spec).surface_forms(tokens, int(g["maxlen"]), spec["pad_token_id"], sp, int(g["padding"]),
                                                    n_threads=threads)
    assert out.dtype == np.int32
    np.testing.assert_array_equal(out, g["matrix"])
    assert n_trunc == int(g["n_truncated"])


@pytest.mark.parametrize("kind", ["unigram", "bpe"])
def test_public_function_matches_hf_loop(kind):
    """get_surface_form_matrix(tokens, maxlen, hn_tokenizer) == the reference's loop over the HF wheel."""
    hn = synthetic.make_hn_tokenizer(kind, 3000, seed=3)
    tokens = synthetic.make_target_tokens(4000, seed=4, specials=("</s>", "<unk>"))
    got, nt = get_surface_form_matrix(tokens, 7, hn, padding=3)
    want, nt2 = ro.surface_form_matrix_hf(tokens, 7, hn, padding=3)
    np.testing.assert_array_equal(got, want)
    assert nt == nt2 and got.shape == (4003, 7)
    assert (got[-3:] == hn.pad_token_id).all()
    assert got[0, 0] == hn.convert_tokens_to_ids("</s>") and (got[0, 1:] == hn.pad_token_id).all()


def test_tokenize_randomised_against_hf_and_oracle():
    from tokenizers import models
    rng = np.random.default_rng(11)
    alphabet = ["a", "b", "c", "Ġ", "é", "ł"]
    for trial in range(25):
        pieces = {"<unk>": 0.0}
        for _ in range(int(rng.integers(5, 40))):
            n = int(rng.integers(1, 5))
            pieces["".join(alphabet[int(i)] for i in rng.integers(0, len(alphabet) - 1, size=n))] = float(-rng.integers(1, 6))
        vocab = list(pieces.items())
        hf = models.Unigram(vocab, unk_id=0, byte_fallback=False)
        nat = NativeTokenizerModel.unigram(vocab, 0)
        orc = ro.UnigramOracle(vocab, 0)
        for _ in range(200):
            s = "".join(alphabet[int(i)] for i in rng.integers(0, len(alphabet), size=int(rng.integers(0, 12))))
            want = [t.id for t in hf.tokenize(s)] if s else []
            assert nat.tokenize(s) == want == orc.tokenize(s), (vocab, s)


def test_bpe_randomised_against_hf():
    from tokenizers import models
    rng = np.random.default_rng(12)
    alphabet = ["a", "b", "c", "d", "é"]
    for trial in range(20):
        vocab = {"<unk>": 0}
        for ch in alphabet[:-1] if trial % 2 else alphabet:
            vocab[ch] = len(vocab)
        merges = []
        syms = [s for s in vocab if s != "<unk>"]
        for _ in range(int(rng.integers(3, 25))):
            a, b = syms[int(rng.integers(0, len(syms)))], syms[int(rng.integers(0, len(syms)))]
            if (a, b) in merges:
                continue
            merges.append((a, b))
            if a + b not in vocab:
                vocab[a + b] = len(vocab)
                syms.append(a + b)
        for fuse in (False, True):
            hf = models.BPE(vocab=vocab, merges=merges, unk_token="<unk>", fuse_unk=fuse)
            nat = NativeTokenizerModel.bpe(vocab, merges, unk_token="<unk>", fuse_unk=fuse)
            for _ in range(150):
                s = "".join(alphabet[int(i)] for i in rng.integers(0, len(alphabet), size=int(rng.integers(0, 14))))
                want = [t.id for t in hf.tokenize(s)] if s else []
                assert nat.tokenize(s) == want, (vocab, merges, fuse, s)


def test_key_error_and_missing_unk():
    hn = synthetic.make_hn_tokenizer("unigram", 600, seed=3)
    with pytest.raises(KeyError):
        get_surface_form_matrix(["ab c"], 7, hn)       # a raw space is not in the byte alphabet (utils.py:675)
    with pytest.raises(KeyError):
        get_surface_form_matrix(["ok", "日本"], 7, hn)
    nat = NativeTokenizerModel.unigram([("a", -1.0), ("b", -2.0)], None)
    assert nat.tokenize("ab") == [0, 1]
    with pytest.raises(Exception):
        nat.tokenize("abz")


def test_empty_and_ragged_inputs():
    hn = synthetic.make_hn_tokenizer("unigram", 600, seed=3)
    out, nt = get_surface_form_matrix([], 7, hn, padding=2)
    assert out.shape == (2, 7) and nt == 0 and (out == hn.pad_token_id).all()
    out, nt = get_surface_form_matrix(["", "a", "a" * 40], 5, hn)
    want, nt2 = ro.surface_form_matrix_hf(["", "a", "a" * 40], 5, hn)
    np.testing.assert_array_equal(out, want)
    assert nt == nt2 == 1


def test_tokenizer_object_input():
    """First argument may be a tokenizer: rows follow convert_ids_to_tokens(range(len(tokenizer))) (utils.py:655-659)."""
    hn = synthetic.make_hn_tokenizer("unigram", 600, seed=3)
    target = synthetic.make_hn_tokenizer("unigram", 500, seed=9)
    got, _ = get_surface_form_matrix(target, 7, hn)
    want, _ = ro.surface_form_matrix_hf(target.convert_ids_to_tokens(range(len(target))), 7, hn)
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("case", INT_CASES)
def test_blob_entry_point_equals_pointer_array(golden_dir, case):
    """zett_surface_forms_blob (one NUL-separated buffer, special tokens matched natively) == zett_surface_forms (char*
    array, per-token special ids) == the reference-minted golden; an embedded NUL falls back to the array path."""
    g = np.load(os.path.join(golden_dir, f"surface_forms_{case}.npz"))
    spec, tokens = json.loads(str(g["spec"])), json.loads(str(g["tokens"]))
    model = native_model(spec)
    out, n_trunc = model.surface_forms(tokens, int(g["maxlen"]), spec["pad_token_id"], None, int(g["padding"]),
                                       special_tokens=spec["special_tokens"])
    np.testing.assert_array_equal(out, g["matrix"])
    assert n_trunc == int(g["n_truncated"])
    # ragged edges: no tokens, one empty token, empty tokens at both ends, no special tokens at all
    for toks in ([], [""], ["", tokens[1], ""], tokens[:1]):
        sp = np.array([spec["special_tokens"].get(t, -1) for t in toks], dtype=np.int32)
        a, na = model.surface_forms(toks, 5, spec["pad_token_id"], sp)
        b, nb = model.surface_forms(toks, 5, spec["pad_token_id"], special_tokens=spec["special_tokens"])
        c, nc = model.surface_forms(toks, 5, spec["pad_token_id"], special_tokens={})
        np.testing.assert_array_equal(a, b)
        assert na == nb and a.shape == (len(toks), 5)
        if not any(t in spec["special_tokens"] for t in toks):
            np.testing.assert_array_equal(a, c)


def _byte_tokens():
    return ["<0x%02X>" % b for b in range(256)]


def test_option_branches_against_hf():
    """The model options the shipped hn tokenizers do not use, each against the HF wheel on randomised input: Unigram and
    BPE byte_fallback (unknown chars become <0xXX> ids), BPE continuing_subword_prefix / end_of_word_suffix, BPE without an
    unk token (unknown chars are dropped), ignore_merges."""
    from tokenizers import models
    rng = np.random.default_rng(21)
    alphabet = ["a", "b", "c", "é", "ł", "z"]   # "z" and sometimes "ł" are outside the vocabularies below

    def strings(k=200, maxlen=10):
        for _ in range(k):
            yield "".join(alphabet[int(i)] for i in rng.integers(0, len(alphabet), size=int(rng.integers(0, maxlen))))

    # --- Unigram byte_fallback
    for trial in range(6):
        vocab = [("<unk>", 0.0)] + [(t, 0.0) for t in _byte_tokens()]
        for _ in range(int(rng.integers(4, 25))):
            n = int(rng.integers(1, 4))
            p = "".join(alphabet[int(i)] for i in rng.integers(0, 4, size=n))
            if p not in dict(vocab):
                vocab.append((p, float(-rng.integers(1, 6))))
        hf = models.Unigram(vocab, unk_id=0, byte_fallback=True)
        nat = NativeTokenizerModel.unigram(vocab, 0, byte_fallback=True)
        for s in strings():
            want = [t.id for t in hf.tokenize(s)] if s else []
            assert nat.tokenize(s) == want, ("unigram byte_fallback", vocab[257:], s)

    # --- BPE: byte_fallback / prefix + suffix / no unk / ignore_merges
    def random_bpe(chars, extra_tokens=(), prefix="", suffix=""):
        vocab = {}
        for t in extra_tokens:
            vocab[t] = len(vocab)
        forms = set()
        for ch in chars:
            forms.update({ch, prefix + ch, ch + suffix, prefix + ch + suffix})
        for f in sorted(forms):
            if f not in vocab:
                vocab[f] = len(vocab)
        syms = [s for s in vocab if s not in extra_tokens]
        merges = []
        for _ in range(int(rng.integers(3, 20))):
            a, b = syms[int(rng.integers(0, len(syms)))], syms[int(rng.integers(0, len(syms)))]
            if prefix:
                # HF builds the merged token as a + b[len(prefix):] unconditionally: the right part of a merge is a
                # continuing symbol, and a symbol carrying the end-of-word suffix cannot be a left part
                if not b.startswith(prefix) or (suffix and a.endswith(suffix)):
                    continue
                merged = a + b[len(prefix):]
            else:
                merged = a + b
            if (a, b) in merges:
                continue
            merges.append((a, b))
            if merged not in vocab:
                vocab[merged] = len(vocab)
                syms.append(merged)
        return vocab, merges

    for trial in range(6):
        vocab, merges = random_bpe(alphabet[:4], extra_tokens=["<unk>"] + _byte_tokens())
        for fuse in (False, True):
            hf = models.BPE(vocab=vocab, merges=merges, unk_token="<unk>", fuse_unk=fuse, byte_fallback=True)
            nat = NativeTokenizerModel.bpe(vocab, merges, unk_token="<unk>", fuse_unk=fuse, byte_fallback=True)
            for s in strings():
                want = [t.id for t in hf.tokenize(s)] if s else []
                assert nat.tokenize(s) == want, ("bpe byte_fallback", fuse, merges, s)
    for trial in range(6):
        vocab, merges = random_bpe(alphabet[:4], extra_tokens=["<unk>"], prefix="##", suffix="</w>")
        hf = models.BPE(vocab=vocab, merges=merges, unk_token="<unk>", continuing_subword_prefix="##", end_of_word_suffix="</w>")
        nat = NativeTokenizerModel.bpe(vocab, merges, unk_token="<unk>", continuing_subword_prefix="##", end_of_word_suffix="</w>")
        for s in strings():
            want = [t.id for t in hf.tokenize(s)] if s else []
            assert nat.tokenize(s) == want, ("bpe prefix/suffix", merges, s)
    for trial in range(6):
        vocab, merges = random_bpe(alphabet[:4])
        hf = models.BPE(vocab=vocab, merges=merges)                       # no unk token: unknown chars vanish
        nat = NativeTokenizerModel.bpe(vocab, merges)
        hf_im = models.BPE(vocab=vocab, merges=merges, ignore_merges=True)
        nat_im = NativeTokenizerModel.bpe(vocab, merges, ignore_merges=True)
        for s in strings():
            want = [t.id for t in hf.tokenize(s)] if s else []
            assert nat.tokenize(s) == want, ("bpe no unk", merges, s)
            want = [t.id for t in hf_im.tokenize(s)] if s else []
            assert nat_im.tokenize(s) == want, ("bpe ignore_merges", merges, s)
