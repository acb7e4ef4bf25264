"""``TokenizerSampler`` -- the reference's one native component (rust_utils/src/lib.rs:20-66), over libzett_b200.so.

Same constructor and call as the PyO3 class the training collator uses (zett/collator.py:341-452):

    sampler = TokenizerSampler()
    pieces = sampler.sample_tokenizer(text_counts, seed_size, max_length, stride=1, noise_std=0.0, pop_prev=True,
                                      push_current=True)              # -> [(piece, log_prob), ...]

plus ``noise_seed`` (the reference's generator is unseeded; nothing is drawn at ``noise_std == 0``).  Host code only.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Tuple

from . import _lib


class TokenizerSampler:
    def __init__(self):
        self.lib = _lib.load()
        self.handle = ctypes.c_void_p()
        _lib.check(self.lib.zett_sampler_create(ctypes.byref(self.handle)))

    def sample_tokenizer(self, map: Dict[str, int], seed_size: int, max_length: int, stride: Optional[int] = None,
                         noise_std: Optional[float] = None, pop_prev: Optional[bool] = None, push_current: Optional[bool] = None,
                         noise_seed: int = 0) -> List[Tuple[str, float]]:
        texts = list(map)
        if any("\0" in t for t in texts):
            raise ValueError("texts must not contain NUL characters")
        blob = ("\0".join(texts) + "\0").encode("utf-8") if texts else b""
        counts = (ctypes.c_uint32 * max(len(texts), 1))(*[int(map[t]) for t in texts])
        out_blob, out_scores = ctypes.c_void_p(), ctypes.c_void_p()
        blob_bytes, n = ctypes.c_int64(0), ctypes.c_int64(0)
        _lib.check(self.lib.zett_sampler_sample(
            self.handle, blob, len(blob), counts, len(texts), int(seed_size), int(max_length), 1 if stride is None else int(stride),
            0.0 if noise_std is None else float(noise_std), int(noise_seed), int(True if pop_prev is None else pop_prev),
            int(True if push_current is None else push_current), ctypes.byref(out_blob), ctypes.byref(blob_bytes),
            ctypes.byref(out_scores), ctypes.byref(n)))
        try:
            raw = ctypes.string_at(out_blob.value, blob_bytes.value)
            pieces = raw.split(b"\0")[: n.value]
            scores = (ctypes.c_double * max(n.value, 1)).from_address(out_scores.value)
            return [(p.decode("utf-8"), float(scores[i])) for i, p in enumerate(pieces)]
        finally:
            self.lib.zett_sampler_free(out_blob)
            self.lib.zett_sampler_free(out_scores)

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.zett_sampler_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
