"""Checkpoint ingestion without JAX/Flax: read a hypernet's ``flax_model.msgpack`` (what the reference's transfer driver
loads, scripts/transfer.py:145-151) and map it onto the PyTorch ``state_dict`` names the kernels are fed with
(the mapping of scripts/convert_to_pt.py:35-45 + transformers' generic Flax->PyTorch rules).

Format (``flax.serialization.msgpack_serialize``): a msgpack map whose leaves are ext type 1 = msgpack-packed
``(shape, dtype name, C-order bytes)``; arrays above 2**30 bytes are stored as
``{"__msgpack_chunked_array__": True, "shape": {"0": ..}, "chunks": {"0": ndarray, ..}}``.  ``flax`` is not installed in
this image, so the reader is pinned only by a writer of the same format in the tests ("parity unpinned" for
this row).
"""
from __future__ import annotations

import os
import re
from typing import Dict

import msgpack
import numpy as np

EXT_NDARRAY, EXT_COMPLEX, EXT_NPSCALAR = 1, 2, 3


def _array_from_ext(data: bytes) -> np.ndarray:
    shape, dtype_name, buf = msgpack.unpackb(data, raw=True)
    name = dtype_name.decode() if isinstance(dtype_name, bytes) else dtype_name
    if name == "bfloat16":  # numpy has no bfloat16: widen to float32 (exact)
        u = np.frombuffer(buf, dtype=np.uint16).astype(np.uint32) << 16
        return u.view(np.float32).reshape(shape)
    return np.frombuffer(buf, dtype=np.dtype(name)).reshape(shape)


def _ext_hook(code: int, data: bytes):
    if code in (EXT_NDARRAY, EXT_NPSCALAR):
        return _array_from_ext(data)
    if code == EXT_COMPLEX:
        re_, im = msgpack.unpackb(data)
        return complex(re_, im)
    return msgpack.ExtType(code, data)


def _unchunk(tree):
    if isinstance(tree, dict):
        if tree.get("__msgpack_chunked_array__"):
            shape = tuple(tree["shape"][str(i)] for i in range(len(tree["shape"])))
            chunks = [tree["chunks"][str(i)] for i in range(len(tree["chunks"]))]
            return np.concatenate(chunks).reshape(shape)
        return {k: _unchunk(v) for k, v in tree.items()}
    return tree


def read_flax_msgpack(path_or_bytes) -> dict:
    """Nested ``{name: {...: ndarray}}`` parameter tree of a ``flax_model.msgpack``."""
    if isinstance(path_or_bytes, (bytes, bytearray)):
        raw = bytes(path_or_bytes)
    else:
        with open(path_or_bytes, "rb") as f:
            raw = f.read()
    return _unchunk(msgpack.unpackb(raw, ext_hook=_ext_hook, raw=False, strict_map_key=False))


def _ndarray_to_ext(a: np.ndarray) -> msgpack.ExtType:
    a = np.ascontiguousarray(a)
    return msgpack.ExtType(EXT_NDARRAY, msgpack.packb((list(a.shape), a.dtype.name, a.tobytes()), use_bin_type=True))


def msgpack_serialize(tree) -> bytes:
    """``flax.serialization.msgpack_serialize`` for a numpy array or a nested dict of them (arrays below the 2**30-byte
    chunking threshold): what the reference writes as ``bias.msgpack`` (scripts/transfer.py:305-310)."""
    def enc(x):
        if isinstance(x, dict):
            return {k: enc(v) for k, v in x.items()}
        if isinstance(x, np.generic):
            return msgpack.ExtType(EXT_NPSCALAR, msgpack.packb((list(np.asarray(x).shape), np.asarray(x).dtype.name, np.asarray(x).tobytes()),
                                                               use_bin_type=True))
        a = np.asarray(x)
        if a.nbytes >= 2 ** 30:
            raise ValueError("arrays of 1 GiB and more need the chunked layout, which this writer does not produce")
        return _ndarray_to_ext(a)
    return msgpack.packb(enc(tree), use_bin_type=True)


def _flatten(tree, prefix=()):
    for k, v in tree.items():
        if isinstance(v, dict):
            yield from _flatten(v, prefix + (str(k),))
        else:
            yield prefix + (str(k),), v


def flax_params_to_state_dict(params: dict) -> Dict[str, np.ndarray]:
    """Flax parameter tree of ``zett.model.Hypernet`` -> reference PyTorch ``state_dict`` (fp32 numpy arrays).

    ``layers_N`` -> ``N``; ``kernel`` -> ``weight`` (transposed); ``scale`` / ``embedding`` -> ``weight``;
    ``model.embeddings.lang_embedding.embedding`` -> ``lang_embeddings.weight``; Rescaler ``w`` / ``b`` keep their names."""
    if set(params) == {"params"}:
        params = params["params"]
    out: Dict[str, np.ndarray] = {}
    for path, value in _flatten(params):
        arr = np.asarray(value)
        parts = [re.sub(r"^layers_(\d+)$", r"\1", p) for p in path]
        leaf = parts[-1]
        if parts[:3] == ["model", "embeddings", "lang_embedding"]:
            out["lang_embeddings.weight"] = arr.astype(np.float32)
            continue
        if leaf == "kernel":
            parts[-1] = "weight"
            arr = arr.T
        elif leaf in ("scale", "embedding"):
            parts[-1] = "weight"
        out[".".join(parts)] = np.ascontiguousarray(arr, dtype=np.float32)
    return out


def load_flax_hypernet(checkpoint_path: str):
    """``ZettHypernet`` from a reference training checkpoint directory (``config.json`` + ``flax_model.msgpack``)."""
    import torch
    from .config import ZettHypernetConfig
    from .modeling_hypernet import ZettHypernet

    config = ZettHypernetConfig.from_pretrained(checkpoint_path)
    model = ZettHypernet(config)
    sd = flax_params_to_state_dict(read_flax_msgpack(os.path.join(checkpoint_path, "flax_model.msgpack")))
    # The Flax hypernet is built with vocab_size = 1, so a checkpoint may carry a [1, H] word-embedding table; it is never
    # read (the forward feeds inputs_embeds) and the reference's converter tolerates the mismatch
    # (scripts/convert_to_pt.py:35-45 goes through load_flax_weights_in_pytorch_model): drop it before checking.
    sd = {k: v for k, v in sd.items() if "word_embeddings" not in k}
    own = model.state_dict()
    unexpected = sorted(set(sd) - set(own))
    missing = sorted(k for k in set(own) - set(sd) if "word_embeddings" not in k)  # never read (inputs_embeds path)
    if unexpected or missing:
        raise ValueError(f"checkpoint does not match the config: missing={missing} unexpected={unexpected}")
    for k, v in sd.items():
        if tuple(own[k].shape) != tuple(v.shape):
            raise ValueError(f"{k}: checkpoint shape {v.shape} != expected {tuple(own[k].shape)}")
    model.load_state_dict({k: torch.from_numpy(np.array(v, copy=True)) for k, v in sd.items()}, strict=False)
    return model
