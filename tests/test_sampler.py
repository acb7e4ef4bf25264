"""Native TokenizerSampler (C ABI, host code) against the Python restatement of rust_utils/src/lib.rs:69-250, which takes
its pre-tokenisation from the HF wheel -- the same crate the reference links.  Exact at noise_std = 0."""
import numpy as np
import pytest

from oracle.sampler_oracle import TokenizerSamplerOracle, pretokenize, substring_scores
from zett_b200.sampler import TokenizerSampler

TEXTS = [
    "hello world", "Hello, World!  It's 12:30 -- don't panic", "héllo wörld's naïve café", "日本語のテキスト と 한국어 텍스트",
    "tabs\tand\nnewlines\r\n  and   spaces   ", "x²+y³=z⁴ ½ Ⅻ ٣٤", "emoji 🙂🙃 and symbols ©®™ §¶", "'s't're've'm'll'd 'S 'x", "   ",
    "a", "", "we'll they've I'm he'd can't", "mixed123abc 4five6", "Ünïcödé ŁÓDŹ straße ﬁ", "trailing space ", " leading", " nbsp em　cjk ",
]


def random_texts(rng, n):
    pools = ["abcdefghij", "ABC", "0123", " ", "  ", "\n", "\t", "'", ".,!?-", "éöłß", "日本", "한", "²½", "🙂", " ", "_", "'s", "'ll"]
    out = {}
    for _ in range(n):
        k = int(rng.integers(0, 14))
        s = "".join(pools[int(i)] for i in rng.integers(0, len(pools), size=k))
        out[s] = int(rng.integers(1, 50))
    return out


def as_map(pairs):
    return dict(pairs)


def test_scores_and_seed_list_match_the_restatement():
    rng = np.random.default_rng(0)
    texts = {t: int(rng.integers(1, 9)) for t in TEXTS}
    for max_length, stride, seed_size in ((5, 1, 2000), (3, 2, 400), (9, 1, 100000)):
        nat, orc = TokenizerSampler(), TokenizerSamplerOracle()
        got = nat.sample_tokenizer(texts, seed_size, max_length, stride=stride)
        want = orc.sample_tokenizer(texts, seed_size, max_length, stride=stride)
        assert [p for p, _ in got] == [p for p, _ in want]
        np.testing.assert_allclose([s for _, s in got], [s for _, s in want], rtol=1e-15, atol=0)


def test_randomised_batches_and_cache_protocol():
    """Several calls on one sampler: the seed cache (pop_prev / push_current) must evolve like the reference's."""
    rng = np.random.default_rng(1)
    nat, orc = TokenizerSampler(), TokenizerSamplerOracle()
    for step in range(8):
        texts = random_texts(rng, 40)
        pop_prev, push_current = bool(step % 3 != 1), bool(step % 4 != 2)
        got = nat.sample_tokenizer(texts, 3000, 6, stride=1 + step % 2, pop_prev=pop_prev, push_current=push_current)
        want = orc.sample_tokenizer(texts, 3000, 6, stride=1 + step % 2, pop_prev=pop_prev, push_current=push_current)
        assert [p for p, _ in got] == [p for p, _ in want], step
        np.testing.assert_allclose([s for _, s in got], [s for _, s in want], rtol=1e-15, atol=0)
        if not pop_prev:
            assert got == []


def test_structure_of_the_seed_list():
    nat = TokenizerSampler()
    got = nat.sample_tokenizer({"the quick brown fox": 5, "jumps over the lazy dog": 2}, 400, 4)
    pieces = [p for p, _ in got]
    assert len(pieces) == len(set(pieces)) and len(pieces) <= 400 + 1
    assert len(pieces[0]) == 1 and len({s for _, s in got[:256]}) == 1          # 256 byte-level characters at the minimum log-prob
    assert all(s == 0.0 for _, s in got[256:256 + 27])                          # 3 x (max_length - 1) x 3 whitespace runs
    rest = got[256 + 27:]
    assert all(a[1] >= b[1] for a, b in zip(rest, rest[1:]))                    # by decreasing score
    assert "Ġth" in pieces and all(len(p) < 4 for p in pieces[256 + 27:])   # substrings are shorter than max_length


def test_noise_is_seeded():
    nat = TokenizerSampler()
    texts = {"alpha beta gamma delta": 3, "beta gamma": 7}
    a = nat.sample_tokenizer(texts, 500, 5, noise_std=1e-3, push_current=False, noise_seed=4)
    b = nat.sample_tokenizer(texts, 500, 5, noise_std=1e-3, push_current=False, noise_seed=4)
    c = nat.sample_tokenizer(texts, 500, 5, noise_std=1e-3, push_current=False, noise_seed=5)
    assert a == b and a != c
    with pytest.raises(ValueError):
        nat.sample_tokenizer({"a": 1}, 10, 0)


def test_pretokenizer_scanner_on_unicode_samples():
    """The hand-written GPT-2 scanner against the wheel's regex engine, via the scores of length-1 substrings (every byte of
    every pre-token is the start of one) on text covering letters, digits, marks, symbols and white space of many scripts."""
    rng = np.random.default_rng(2)
    cps = [c for c in list(range(0x20, 0x250)) + list(range(0x370, 0x400)) + list(range(0x600, 0x6FF)) + list(range(0x900, 0x97F)) +
           list(range(0x2000, 0x2070)) + list(range(0x2150, 0x2190)) + list(range(0x3000, 0x30FF)) + list(range(0x4E00, 0x4E40)) +
           list(range(0xFF10, 0xFF5A)) + list(range(0x1D7CE, 0x1D7F0)) + [0x9, 0xA, 0xD, 0x85, 0xA0, 0x1F600] if not (0xD800 <= c <= 0xDFFF)]
    for _ in range(30):
        text = "".join(chr(cps[int(i)]) for i in rng.integers(0, len(cps), size=60))
        texts = {text: 1}
        got = as_map(TokenizerSampler().sample_tokenizer(texts, 10 ** 6, 8, push_current=False))
        want = as_map(TokenizerSamplerOracle().sample_tokenizer(texts, 10 ** 6, 8, push_current=False))
        assert got == want, [hex(ord(c)) for c in text]
    assert pretokenize(" a")[0][0] == "Ġa" and substring_scores({"a": 2}, 3)["Ġa"] == 2 * 3 * 2   # listed twice: the duplicate 0
