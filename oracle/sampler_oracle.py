"""Python restatement of ``rust_utils.TokenizerSampler.sample_tokenizer`` (reference rust_utils/src/lib.rs:69-250) --
TEST INFRASTRUCTURE, NOT PRODUCT CODE (same import rule as the rest of oracle/).

The reference builds its pre-tokenizer from the HF ``tokenizers`` crate (Split on the GPT-2 regex, inverted, then
ByteLevel without its own regex; lib.rs:83-93); the same two pre-tokenizers are in the HF wheel, so this restatement calls
them for the splits and their character offsets and restates only the reference's own arithmetic: the cumulative byte
lengths (lib.rs:101-107), the per-pretoken start list (lib.rs:123-130, including its duplicate 0 for the first pretoken
and its end-of-character rather than start-of-character offsets for multi-byte characters), the substring scores
(lib.rs:134-158), the seed cache (lib.rs:163-176, 240-245) and the seed list (lib.rs:178-238).

What the reference leaves to chance -- Gaussian noise from thread_rng (lib.rs:199-200), hash-map iteration order for the
alphabet and for ties -- is pinned here: noise comes from a seeded numpy generator (and is absent at noise_std = 0), the
alphabet is emitted in byte order, ties in the sort are broken by the piece's string.  The native port
(zett_b200/csrc/sampler.cpp) makes the same choices and is compared with this file at noise_std = 0.
"""
from __future__ import annotations

import math
from collections import deque
from typing import Dict, List, Tuple

import numpy as np
from tokenizers import Regex, pre_tokenizers

SPLIT_REGEX = r"'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+"   # lib.rs:27
EXTRA_WHITESPACE = ["Ġ", "Ċ", "ĉ"]                                                               # lib.rs:212


def byte_alphabet() -> List[str]:
    """ByteLevel::alphabet() in byte order (the GPT-2 bytes_to_unicode table)."""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    cs, n = bs[:], 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    table = dict(zip(bs, cs))
    return [chr(table[b]) for b in range(256)]


def pretokenize(sentence: str) -> List[Tuple[str, Tuple[int, int]]]:
    pt = pre_tokenizers.Sequence([pre_tokenizers.Split(Regex(SPLIT_REGEX), "removed", invert=True),
                                  pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)])
    return pt.pre_tokenize_str(sentence)


def substring_scores(texts: Dict[str, int], max_length: int, stride: int = 1) -> Dict[str, int]:
    """lib.rs:95-161: score of every byte-level substring of at most max_length - 1 chars that starts at a listed offset."""
    index: Dict[str, int] = {}
    for sentence, n in texts.items():
        sentence = " " + sentence
        cum, acc = [], 0
        for ch in sentence:
            acc += len(ch.encode("utf-8"))
            cum.append(acc)
        for i, (pretoken, (o0, o1)) in enumerate(pretokenize(sentence)):
            starts = [cum[j] - cum[o0] for j in range(o0, o1)]
            if i == 0:
                starts.insert(0, 0)
            nb = len(pretoken)   # byte-level chars = bytes of the original piece
            for s in starts[::stride]:
                for k in range(1, max_length):
                    if s + k > nb:
                        break
                    token = pretoken[s:s + k]
                    if not token:
                        continue
                    index[token] = index.get(token, 0) + n * len(token.encode("utf-8"))
    return index


class TokenizerSamplerOracle:
    def __init__(self):
        self.seed_cache: deque = deque()

    def sample_tokenizer(self, texts: Dict[str, int], seed_size: int, max_length: int, stride: int = 1, noise_std: float = 0.0,
                         pop_prev: bool = True, push_current: bool = True, noise_seed: int = 0) -> List[Tuple[str, float]]:
        current = substring_scores(texts, max_length, stride)
        maybe_prev = self.seed_cache.pop() if (pop_prev and self.seed_cache) else None
        self.seed_cache.appendleft(current)
        out: List[Tuple[str, float]] = []
        if pop_prev:
            merged: Dict[str, int] = {}
            for m in self.seed_cache:
                for k, v in m.items():
                    merged[k] = merged.get(k, 0) + v
            score_sum = float(sum(merged.values()))
            min_score = float(min(merged.values())) if merged else float(2 ** 32 - 1)
            min_log_prob = math.log(min_score / score_sum) if score_sum > 0 else float("nan")
            out.extend((c, min_log_prob) for c in byte_alphabet())
            rng = np.random.default_rng(noise_seed)
            items = sorted(merged.items())   # deterministic order before the noise is drawn
            noise = rng.normal(0.0, noise_std, size=len(items)) if noise_std > 0 else np.zeros(len(items))
            scored = []
            for (tok, v), eps in zip(items, noise):
                noised = v / score_sum + float(eps)
                scored.append((tok, math.log(noised) if noised > 0 else -100000.0))
            scored.sort(key=lambda kv: (-kv[1], kv[0]))
            for c1 in EXTRA_WHITESPACE:
                for i in range(1, max_length):
                    for c2 in EXTRA_WHITESPACE:
                        out.append((c2 + c1 * i, 0.0))
            for tok, score in scored:
                if len(tok) == 1 or sum(ch in EXTRA_WHITESPACE for ch in tok) >= 2:
                    continue
                out.append((tok, score))
                if len(out) >= seed_size:
                    break
        if not push_current:
            self.seed_cache.popleft()
            if maybe_prev is not None:
                self.seed_cache.append(maybe_prev)
        return out
