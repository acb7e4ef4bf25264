"""Benchmark / test infrastructure, NOT part of the product package.

Seeded synthetic workloads: hypernet configs at the BASELINE shapes, random
weights with the reference's ``state_dict`` names, source-embedding tables, hn tokenizers (Unigram / BPE)
and byte-level target vocabularies.  No real tokenizer or checkpoint is reachable offline (the reference's
``artifacts/tokenizers/*/tokenizer.json`` are Git-LFS pointers), so tests, ``smoke()`` and ``bench.py``
all draw from these generators.  numpy ``default_rng`` keeps the streams identical on every box.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

from zett_b200.config import ZettHypernetConfig, weight_shapes

# --------------------------------------------------------------------------------------------------
# configs (hyper-parameters from the reference's shipped configs)
# --------------------------------------------------------------------------------------------------
_SHAPES = {
    # name: dict(n_embd, hidden, intermediate, heads, lang_id, n_langs, pad, v0, separate, n_extra)
    "tiny": dict(n_embd=64, hn_hidden_size=128, hn_intermediate_size=256, hn_num_attention_heads=None,
                 hn_embed_lang_id=False, n_langs=None, pad_token_id=2, original_vocab_size=300,
                 separate_out_embeddings=True, hn_n_extra_tokens=16),
    "tiny_lang": dict(n_embd=128, hn_hidden_size=128, hn_intermediate_size=256, hn_num_attention_heads=4,
                      hn_embed_lang_id=True, n_langs=5, pad_token_id=1, original_vocab_size=300,
                      separate_out_embeddings=False, hn_n_extra_tokens=0),
    # configs/zeroshot/v7:xlmr:multilingual_long:lw=0.5_26l.json:61,66-67
    "xlmr": dict(n_embd=768, hn_hidden_size=768, hn_intermediate_size=1536, hn_num_attention_heads=None,
                 hn_embed_lang_id=True, n_langs=26, pad_token_id=1, original_vocab_size=250002,
                 separate_out_embeddings=False, hn_n_extra_tokens=256),
    # configs/zeroshot/v7:tinyllama_en+code:lw=0.5_long.json:57-58
    "tinyllama": dict(n_embd=2048, hn_hidden_size=2048, hn_intermediate_size=4096, hn_num_attention_heads=None,
                      hn_embed_lang_id=False, n_langs=None, pad_token_id=2, original_vocab_size=32000,
                      separate_out_embeddings=True, hn_n_extra_tokens=256),
    # configs/zeroshot/v7:mistral7b_en+code:lw=0.5_long.json:58-60
    "mistral": dict(n_embd=4096, hn_hidden_size=4096, hn_intermediate_size=8192, hn_num_attention_heads=32,
                    hn_embed_lang_id=False, n_langs=None, pad_token_id=2, original_vocab_size=32000,
                    separate_out_embeddings=True, hn_n_extra_tokens=256),
}


def make_config(name: str, **overrides) -> ZettHypernetConfig:
    kw = dict(
        hn_model_name_or_path="roberta-base", hn_surface_maxlen=7, hn_n_layers=3,
        hn_rescale_embeddings=True, hn_embed_using_source_embeddings=True, hn_predict_bias=True,
    )
    kw.update(_SHAPES[name])
    kw.update(overrides)
    return ZettHypernetConfig(**kw)


def config_names() -> List[str]:
    return list(_SHAPES)


# --------------------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------------------
def make_weights(cfg: ZettHypernetConfig, seed: int = 0) -> Dict[str, np.ndarray]:
    """Random fp32 weights: Linear W ~ N(0, 1/fan_in), biases ~ N(0, 0.02^2),
    LayerNorm gamma = 1 + 0.1 N(0,1), beta = 0.1 N(0,1), embeddings ~ N(0, 0.02^2),
    scalers w ~ U(0.5, 1.5), b ~ N(0, 0.02^2)."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for name, shape in weight_shapes(cfg).items():
        leaf = name.rsplit(".", 1)[-1]
        if "LayerNorm" in name or ".ln." in name:
            if leaf == "weight":
                w = 1.0 + 0.1 * rng.standard_normal(shape, dtype=np.float32)
            else:
                w = 0.1 * rng.standard_normal(shape, dtype=np.float32)
        elif "scaler" in name:
            if leaf == "w":
                w = rng.uniform(0.5, 1.5, size=shape).astype(np.float32)
            else:
                w = 0.02 * rng.standard_normal(shape, dtype=np.float32)
        elif "embeddings" in name:
            w = 0.02 * rng.standard_normal(shape, dtype=np.float32)
        elif leaf == "bias":
            w = 0.02 * rng.standard_normal(shape, dtype=np.float32)
        else:  # Linear weight [out, in]
            w = rng.standard_normal(shape, dtype=np.float32) / np.float32(np.sqrt(shape[-1]))
        out[name] = np.ascontiguousarray(w, dtype=np.float32)
    return out


def make_source_embeddings(cfg: ZettHypernetConfig, seed: int = 100, n_rows: Optional[int] = None) -> np.ndarray:
    rng = np.random.default_rng(seed)
    v0 = cfg.original_vocab_size if n_rows is None else n_rows
    return (0.02 * rng.standard_normal((v0, cfg.n_in_embd), dtype=np.float32)).astype(np.float32)


def make_random_surface_forms(cfg: ZettHypernetConfig, n_rows: int, seed: int = 7,
                              p_fallback: float = 0.05, full_rows: bool = False) -> np.ndarray:
    """Random int32 [n_rows, L] surface forms with a prefix of ids and a pad tail, fallback ids mixed in,
    one fully padded row and one row whose position 0 is the pad id (edge cases of SURVEY 8a)."""
    rng = np.random.default_rng(seed)
    L = cfg.hn_surface_maxlen
    v0, nx, pad = cfg.original_vocab_size, cfg.hn_n_extra_tokens, cfg.pad_token_id
    ids = rng.integers(0, v0, size=(n_rows, L)).astype(np.int32)
    ids[ids == pad] = pad + 1
    if nx > 0:
        fb = rng.random((n_rows, L)) < p_fallback
        ids[fb] = (v0 + rng.integers(0, nx, size=int(fb.sum()))).astype(np.int32)
    if not full_rows:
        lens = rng.integers(1, L + 1, size=n_rows)
        ids[np.arange(L)[None, :] >= lens[:, None]] = pad
        if n_rows > 3:
            ids[3, :] = pad            # fully masked row (or lang slot only)
        if n_rows > 5:
            ids[5, 0] = pad            # pad in position 0, real ids after it
            ids[5, 1] = pad + 1
    return ids


from zett_b200.byte_alphabet import BYTES_TO_CHARS, CHARS_TO_BYTES  # noqa: E402,F401

SPECIALS = ["<s>", "<pad>", "</s>", "<unk>"]


def _random_pieces(n: int, seed: int) -> List[str]:
    """Distinct random lowercase pieces, length min(16, max(2, Geometric(0.25))), half prefixed with G-dot."""
    rng = np.random.default_rng(seed)
    space = BYTES_TO_CHARS[32]
    seen, out = set(), []
    while len(out) < n:
        m = n - len(out)
        lens = np.minimum(16, np.maximum(2, rng.geometric(0.25, size=m)))
        pref = rng.random(m) < 0.5
        letters = rng.integers(97, 123, size=(m, 16))
        for i in range(m):
            s = (space if pref[i] else "") + "".join(map(chr, letters[i, : lens[i]]))
            if s not in seen:
                seen.add(s)
                out.append(s)
    return out


def make_hn_vocab(n_vocab: int = 32000, seed: int = 1):
    """hn-tokenizer vocabulary: 4 specials + 256 alphabet chars + random pieces; Unigram scores
    -U(4, 11) for pieces and -12 for single chars (specials 0)."""
    rng = np.random.default_rng(seed + 1000)
    alphabet = [BYTES_TO_CHARS[b] for b in range(256)]
    pieces = _random_pieces(n_vocab - len(SPECIALS) - 256, seed)
    vocab = SPECIALS + alphabet + pieces
    scores = np.concatenate([
        np.zeros(len(SPECIALS)), np.full(256, -12.0), -rng.uniform(4.0, 11.0, size=len(pieces))])
    return vocab, scores.astype(np.float64)


def make_bpe_merges(vocab: List[str]) -> Tuple[List[str], List[Tuple[str, str]]]:
    """Derive a BPE merge list by greedy left-to-right pair merges of each multi-char piece; any
    intermediate symbol missing from the vocabulary is appended to it."""
    vocab = list(vocab)
    index = {t: i for i, t in enumerate(vocab)}
    merges, seen = [], set()
    for piece in list(vocab):
        if piece in SPECIALS or len(piece) < 2:
            continue
        left = piece[0]
        for ch in piece[1:]:
            if (left, ch) not in seen:
                seen.add((left, ch))
                merges.append((left, ch))
            left = left + ch
            if left not in index:
                index[left] = len(vocab)
                vocab.append(left)
    return vocab, merges


def make_hn_tokenizer(kind: str = "unigram", n_vocab: int = 32000, seed: int = 1, pad_token: str = "</s>",
                      fit_total: bool = False):
    """A ``PreTrainedTokenizerFast`` wrapping a synthetic Unigram or BPE model, as ``tokenizer_to_use``.

    BPE vocabularies grow by the intermediate merge symbols (about 3.5x); ``fit_total=True`` shrinks the piece list
    until the FINAL vocabulary has at most ``n_vocab`` entries, so that every id is a valid source-embedding row."""
    from tokenizers import Tokenizer, models
    from transformers import PreTrainedTokenizerFast

    vocab, scores = make_hn_vocab(n_vocab, seed)
    if kind == "unigram":
        model = models.Unigram([(t, float(s)) for t, s in zip(vocab, scores)], unk_id=3, byte_fallback=False)
    elif kind == "bpe":
        base = n_vocab
        vocab, merges = make_bpe_merges(vocab)
        while fit_total and len(vocab) > n_vocab:
            base = max(300, int(base * min(0.97, n_vocab / len(vocab))))
            vocab, merges = make_bpe_merges(make_hn_vocab(base, seed)[0])
        model = models.BPE(vocab={t: i for i, t in enumerate(vocab)}, merges=merges, unk_token="<unk>")
    else:
        raise ValueError(kind)
    return PreTrainedTokenizerFast(tokenizer_object=Tokenizer(model), bos_token="<s>", pad_token=pad_token,
                                   eos_token="</s>", unk_token="<unk>")


def make_target_tokens(n_tokens: int, seed: int = 2, specials: Tuple[str, ...] = ("</s>",)) -> List[str]:
    """'GPT2-style' byte-level target vocabulary: special token(s) + 256 chars + random pieces."""
    alphabet = [BYTES_TO_CHARS[b] for b in range(256)]
    base = list(specials) + alphabet
    if n_tokens <= len(base):
        return base[:n_tokens]
    return base + _random_pieces(n_tokens - len(base), seed)


def length_histogram(surface_forms: np.ndarray, pad_token_id: int) -> List[int]:
    n = (np.asarray(surface_forms) != pad_token_id).sum(axis=1)
    return np.bincount(n, minlength=surface_forms.shape[1] + 1).tolist()


def make_concat_tokens(n_tokens: int, hn_tokenizer, seed: int = 3, specials: Tuple[str, ...] = ("</s>",),
                       length_p=(0.20, 0.40, 0.22, 0.10, 0.05, 0.02, 0.01)) -> List[str]:
    """A second target vocabulary (bench.py ``extra``): every token is a concatenation of 1..7 pieces DRAWN FROM THE hn
    TOKENIZER'S OWN VOCABULARY, so its retokenisation uses ids spread over the whole source table instead of the few short
    pieces random strings fall apart into -- the de-duplication of the input projection / first encoder layer finds far
    less to share.  ``length_p`` follows the non-pad length histogram of the default vocabulary."""
    rng = np.random.default_rng(seed)
    alphabet = [BYTES_TO_CHARS[b] for b in range(256)]
    special = set(hn_tokenizer.all_special_tokens)
    model = hn_tokenizer.backend_tokenizer.model
    pieces = sorted(t for t in hn_tokenizer.get_vocab() if len(t) >= 3 and t not in special and len(model.tokenize(t)) == 1)
    out = list(specials) + alphabet
    seen = set(out)
    p = np.asarray(length_p, dtype=np.float64)
    p = p / p.sum()
    while len(out) < n_tokens:
        m = n_tokens - len(out)
        ks = rng.choice(np.arange(1, len(p) + 1), size=m, p=p)
        idx = rng.integers(0, len(pieces), size=(m, len(p)))
        for i in range(m):
            s = "".join(pieces[j] for j in idx[i, : ks[i]])
            if s not in seen:
                seen.add(s)
                out.append(s)
    return out[:n_tokens]
