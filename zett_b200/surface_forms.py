"""``get_surface_form_matrix`` -- the reference's signature (zett/utils.py:651-689) over the native retokenizer.

The reference walks the vocabulary in a Python loop and calls HF ``tokenizers``' ``Model.tokenize`` once per token;
here the hn tokenizer's model (Unigram or BPE) is exported once into ``libzett_b200.so`` (``zett_tok_create_*``) and
the whole vocabulary is retokenised by ``zett_surface_forms`` on all host cores.  Results are bit-identical
(int32 matrix and truncation count); the same exceptions are raised (``KeyError`` for a char outside the byte
alphabet).  ``ByT5Tokenizer`` hn tokenizers are not supported (no shipped config uses one).
"""
from __future__ import annotations

import ctypes
import json
import weakref
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib


def default_threads() -> int:
    """Worker threads of the retokenizer when the caller does not say: all cores, divided among the processes of a
    one-process-per-GPU launch (LOCAL_WORLD_SIZE / WORLD_SIZE as torchrun sets them) so that eight ranks do not each start a
    full-machine pool on the same host."""
    import os
    cores = os.cpu_count() or 1
    try:
        local = int(os.environ.get("LOCAL_WORLD_SIZE") or os.environ.get("WORLD_SIZE") or 1)
    except ValueError:
        local = 1
    return max(1, cores // max(1, local))


class NativeTokenizerModel:
    """A ``zett_tok`` handle: the Unigram / BPE model of an hn tokenizer."""

    def __init__(self, handle: ctypes.c_void_p, kind: str):
        self.handle = handle
        self.kind = kind
        self.lib = _lib.load()

    @classmethod
    def unigram(cls, vocab: Sequence[Tuple[str, float]], unk_id: Optional[int], byte_fallback: bool = False):
        lib = _lib.load()
        n = len(vocab)
        pieces = (ctypes.c_char_p * n)(*[p.encode("utf-8") for p, _ in vocab])
        scores = (ctypes.c_double * n)(*[float(s) for _, s in vocab])
        h = ctypes.c_void_p()
        _lib.check(lib.zett_tok_create_unigram(pieces, scores, n, -1 if unk_id is None else int(unk_id),
                                               int(bool(byte_fallback)), ctypes.byref(h)))
        return cls(h, "unigram")

    @classmethod
    def bpe(cls, vocab: Dict[str, int], merges: Sequence[Tuple[str, str]], unk_token: Optional[str] = None,
            continuing_subword_prefix: Optional[str] = None, end_of_word_suffix: Optional[str] = None,
            fuse_unk: bool = False, byte_fallback: bool = False, ignore_merges: bool = False):
        lib = _lib.load()
        n = max(vocab.values()) + 1 if vocab else 0
        by_id: List[bytes] = [b"\xff\xfe<unused>"] * n  # ids absent from the vocab can never match a UTF-8 string
        for t, i in vocab.items():
            by_id[i] = t.encode("utf-8")
        arr = (ctypes.c_char_p * n)(*by_id)
        m = len(merges)
        flat = np.empty((m, 2), dtype=np.int32)
        for r, (a, b) in enumerate(merges):
            flat[r, 0], flat[r, 1] = vocab[a], vocab[b]
        h = ctypes.c_void_p()
        unk_id = vocab[unk_token] if unk_token is not None else -1
        _lib.check(lib.zett_tok_create_bpe(
            arr, n, flat.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), m, unk_id,
            continuing_subword_prefix.encode("utf-8") if continuing_subword_prefix else None,
            end_of_word_suffix.encode("utf-8") if end_of_word_suffix else None,
            int(bool(fuse_unk)), int(bool(byte_fallback)), int(bool(ignore_merges)), ctypes.byref(h)))
        return cls(h, "bpe")

    @classmethod
    def from_hf(cls, tokenizer):
        """Export the model of a HF fast tokenizer (``PreTrainedTokenizerFast`` or ``tokenizers.Tokenizer``)."""
        backend = getattr(tokenizer, "_tokenizer", None) or getattr(tokenizer, "backend_tokenizer", None) or tokenizer
        spec = json.loads(backend.to_str())["model"]
        kind = spec.get("type")
        if kind == "Unigram":
            return cls.unigram([(p, s) for p, s in spec["vocab"]], spec.get("unk_id"), spec.get("byte_fallback", False))
        if kind == "BPE":
            merges = [tuple(m.split(" ", 1)) if isinstance(m, str) else tuple(m) for m in spec["merges"]]
            if spec.get("dropout"):
                raise NotImplementedError("BPE dropout is a training-time option")
            return cls.bpe(spec["vocab"], merges, spec.get("unk_token"), spec.get("continuing_subword_prefix"),
                           spec.get("end_of_word_suffix"), spec.get("fuse_unk", False), spec.get("byte_fallback", False),
                           spec.get("ignore_merges", False))
        raise NotImplementedError("hn tokenizer model %r is not supported (Unigram and BPE are)" % kind)

    def tokenize(self, token: str) -> List[int]:
        cap = 4 * len(token.encode("utf-8")) + 8
        buf = (ctypes.c_int32 * cap)()
        n = _lib.check(self.lib.zett_tok_tokenize(self.handle, token.encode("utf-8"), buf, cap))
        return list(buf[:n])

    def surface_forms(self, tokens: Sequence[str], maxlen: int, pad_id: int, special_ids: Optional[np.ndarray] = None,
                      padding: int = 0, n_threads: int = 0, special_tokens: Optional[Dict[str, int]] = None):
        """``special_ids`` (per token, -1 = ordinary) or ``special_tokens`` ({token: id}, matched natively) mark the hn
        tokenizer's special tokens.  The vocabulary crosses the ABI as ONE NUL-separated buffer: marshalling 50k strings
        into a ``char*`` array costs twice the retokenisation itself."""
        v = len(tokens)
        if n_threads <= 0:
            n_threads = default_threads()
        out = np.empty((v + padding, maxlen), dtype=np.int32)
        out_p = out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
        n_trunc = ctypes.c_int64(0)
        blob = "\0".join(tokens).encode("utf-8") if special_ids is None else None
        if blob is not None and blob.count(b"\0") == max(v - 1, 0):  # no token holds a NUL itself
            sp_tokens = list(special_tokens or {})
            sp_blob = "\0".join(sp_tokens).encode("utf-8")
            sp_ids = (ctypes.c_int32 * max(len(sp_tokens), 1))(*[int(special_tokens[t]) for t in sp_tokens])
            _lib.check(self.lib.zett_surface_forms_blob(self.handle, blob, len(blob), v, sp_blob, len(sp_blob), sp_ids,
                                                        len(sp_tokens), maxlen, pad_id, padding, out_p, ctypes.byref(n_trunc),
                                                        n_threads))
            return out, int(n_trunc.value)
        if special_ids is None and special_tokens:
            special_ids = np.fromiter((special_tokens.get(t, -1) for t in tokens), dtype=np.int32, count=v)
        arr = (ctypes.c_char_p * max(v, 1))(*[t.encode("utf-8") for t in tokens])
        sp = None
        if special_ids is not None:
            special_ids = np.ascontiguousarray(special_ids, dtype=np.int32)
            sp = special_ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
        _lib.check(self.lib.zett_surface_forms(self.handle, arr, v, sp, maxlen, pad_id, padding, out_p, ctypes.byref(n_trunc),
                                               n_threads))
        return out, int(n_trunc.value)

    def __del__(self):
        try:
            if self.handle is not None and self.handle.value:
                self.lib.zett_tok_destroy(self.handle)
                self.handle = None
        except Exception:  # noqa: BLE001
            pass


_models: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def native_model_for(tokenizer_to_use) -> NativeTokenizerModel:
    try:
        m = _models.get(tokenizer_to_use)
    except TypeError:
        m = None
    if m is None:
        m = NativeTokenizerModel.from_hf(tokenizer_to_use)
        try:
            _models[tokenizer_to_use] = m
        except TypeError:
            pass
    return m


def get_surface_form_matrix(tokenizer_or_tokens, maxlen, tokenizer_to_use=None, padding=0, verbose=False,
                            n_threads: int = 0):
    """Same contract as the reference (zett/utils.py:651-689): returns ``(int32[V + padding, maxlen], n_truncated)``."""
    if isinstance(tokenizer_or_tokens, list):  # tokens are expected to be byte encoded
        tokens = tokenizer_or_tokens
    else:
        tokenizer = tokenizer_or_tokens
        tokens = tokenizer.convert_ids_to_tokens(range(len(tokenizer)))
    if tokenizer_to_use is None:
        # the reference dereferences tokenizer_to_use unconditionally inside its loop (utils.py:671)
        raise AttributeError("'NoneType' object has no attribute 'all_special_tokens'")
    if type(tokenizer_to_use).__name__ == "ByT5Tokenizer":
        raise NotImplementedError("ByT5 hn tokenizers are not supported by the native retokenizer")
    special_map = {}
    for t in tokenizer_to_use.all_special_tokens:  # first occurrence wins, like `token in all_special_tokens`
        special_map.setdefault(t, int(tokenizer_to_use.convert_tokens_to_ids(t)))
    model = native_model_for(tokenizer_to_use)
    return model.surface_forms(tokens, int(maxlen), int(tokenizer_to_use.pad_token_id), None, int(padding), n_threads,
                               special_tokens=special_map)
