"""``predict`` / ``batched_inference`` -- the inference driver calls of the reference (scripts/transfer.py:54-124,
221-234) over the B200 hypernet.

The reference closes over module-level state (``hypernet_params``, ``source_embeddings_stacked``, ``lang_index``,
``embedding_path_out``, ``bias_path``); here ``make_predict`` builds the same closure explicitly and
``batched_inference`` takes it as an argument.  Semantics are unchanged: rows are shuffled, split in ``batch_size``
chunks, the last chunk is padded by repeating row 0, every chunk's predictions are scatter-added into fp32 host
arrays and the padding is dropped.  ``predict_whole_vocab`` is the B200-first alternative: one call, the library
pipelines its own passes, one device->host copy at the end.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Callable, Optional

import numpy as np
import torch


def default_args(batch_size: int = 16384, sample_batches: bool = False, min_k: int = 1, n_samples: int = 1):
    """The fields of the reference's ``Args`` that ``batched_inference`` reads (scripts/transfer.py:30-51)."""
    return SimpleNamespace(batch_size=batch_size, sample_batches=sample_batches, min_k=min_k, n_samples=n_samples)


def make_predict(hypernet, source_embeddings_stacked, lang_index=None) -> Callable:
    """``predict(target_surface_form_matrix, target_priors)`` of scripts/transfer.py:221-234.

    ``source_embeddings_stacked`` is placed on the hypernet's CUDA device once (the reference's jit closure constant);
    each call copies the surface forms host->device through pinned memory and the three results device->host."""
    nat = hypernet.native()
    device = nat.device
    src = torch.as_tensor(source_embeddings_stacked).to(device=device, dtype=torch.float32).contiguous()
    lang = None if lang_index is None else int(lang_index)

    def predict(target_surface_form_matrix, target_priors=None):
        sf = torch.as_tensor(np.ascontiguousarray(target_surface_form_matrix, dtype=np.int32))
        sf_dev = sf.pin_memory().to(device, non_blocking=True)
        # target_priors only matter for hn_embed_target_priors hypernets, which the PyTorch reference rejects
        pred_in, pred_out, pred_bias = hypernet(sf_dev, source_embeddings=src, lang_index=lang)
        outs = []
        for t in (pred_in, pred_out, pred_bias):
            if t is None:
                outs.append(None)
                continue
            host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            host.copy_(t, non_blocking=True)
            outs.append(host)
        torch.cuda.current_stream(device).synchronize()
        return tuple(None if h is None else h.numpy() for h in outs)

    return predict


def batched_inference(target_surface_form_matrix, target_priors, config, args, predict: Callable,
                      embedding_path_out: Optional[object] = True, bias_path: Optional[object] = True, rng=None):
    """scripts/transfer.py:54-124.  ``embedding_path_out`` / ``bias_path`` keep the reference's meaning: ``None``
    disables the corresponding output array."""
    original_length = len(target_surface_form_matrix)
    if getattr(args, "sample_batches", False):
        # only meaningful for inter-token-attention hypernets, which the PyTorch reference does not support
        raise NotImplementedError("sample_batches")
    rng = np.random if rng is None else rng
    shuffled_indices = rng.permutation(original_length)
    total_length = math.ceil(original_length / args.batch_size) * args.batch_size
    padded = np.pad(shuffled_indices, (0, total_length - original_length))
    indices = np.array_split(padded, total_length // args.batch_size)
    empty_in_last_batch = total_length - original_length

    hidden = getattr(config, "hidden_size", None) or config.n_embd
    predicted_embeddings_in = np.zeros((original_length, hidden), dtype=np.float32)
    predicted_embeddings_out = np.zeros((original_length, hidden), dtype=np.float32) if embedding_path_out is not None else None
    predicted_bias = np.zeros(original_length, dtype=np.float32) if bias_path is not None else None
    if target_priors is None:
        target_priors = np.zeros(original_length, dtype=np.float32)

    for i, batch_indices in enumerate(indices):
        last_batch = i == len(indices) - 1
        in_b, out_b, bias_b = predict(target_surface_form_matrix[batch_indices], target_priors[batch_indices])
        if last_batch and empty_in_last_batch > 0:
            batch_indices = batch_indices[:-empty_in_last_batch]
            in_b = in_b[:-empty_in_last_batch]
            if predicted_embeddings_out is not None:
                out_b = out_b[:-empty_in_last_batch]
            if predicted_bias is not None:
                bias_b = bias_b[:-empty_in_last_batch]
        predicted_embeddings_in[batch_indices] += in_b
        if predicted_embeddings_out is not None:
            predicted_embeddings_out[batch_indices] += out_b
        if predicted_bias is not None:
            predicted_bias[batch_indices] += bias_b
    return predicted_embeddings_in, predicted_embeddings_out, predicted_bias


def predict_whole_vocab(hypernet, target_surface_form_matrix, source_embeddings_dev, lang_index=None):
    """One call for the whole vocabulary: pinned H2D of the int32 matrix, all passes queued back to back on the
    current stream, pinned D2H of the results.  Returns numpy arrays ``(in, out | None, bias)``."""
    return make_predict(hypernet, source_embeddings_dev, lang_index)(target_surface_form_matrix)
