"""Multi-threaded CPU restatement of the hypernetwork forward on torch CPU tensors  --  TEST / BASELINE
INFRASTRUCTURE, NOT PRODUCT CODE (same import rule as oracle/hypernet_oracle.py).

Why a second restatement: the reference's own CPU execution of this path is ``hf_hypernet.ZettHypernet`` on ATen CPU
kernels (threaded GEMM, vectorised GELU / LayerNorm).  numpy runs the element-wise half single-threaded and is several
times slower, which would flatter the GPU/CPU ratio.  This module issues the same ATen operations the reference issues
(``F.linear``, ``F.gelu``, ``F.layer_norm``, eager-attention matmul/softmax -- modeling_hypernet.py:156-267 and HF
``RobertaModel``) so ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg times what the reference would cost on
the same host cores.  ``/root/reference`` itself cannot travel to the GPU box.  Pinned by tests/test_oracle_float.py
against the goldens minted from the reference.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F


def _get(cfg, name, default=None):
    return cfg.get(name, default) if isinstance(cfg, dict) else getattr(cfg, name, default)


def to_torch(weights: Dict[str, np.ndarray], device="cpu") -> Dict[str, torch.Tensor]:
    return {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).to(device) for k, v in weights.items()}


def _projector(x, W, p):
    h = F.gelu(F.linear(x, W[p + "dense1.weight"], W[p + "dense1.bias"]), approximate="tanh")
    h = F.gelu(F.linear(h, W[p + "dense2.weight"], W[p + "dense2.bias"]), approximate="tanh")
    return F.layer_norm(h + x, (x.shape[-1],), W[p + "ln.weight"], W[p + "ln.bias"], 1e-6)


@torch.no_grad()
def hypernet_forward(cfg, W: Dict[str, torch.Tensor], target_surface_forms, source_embeddings: torch.Tensor,
                     lang_index: Optional[int] = None):
    dev = source_embeddings.device
    ids = torch.as_tensor(np.asarray(target_surface_forms)).long().to(dev)
    v0 = int(_get(cfg, "original_vocab_size"))
    pad = int(_get(cfg, "pad_token_id"))
    H = int(_get(cfg, "hn_hidden_size"))
    D = int(_get(cfg, "n_embd"))
    n_layers = int(_get(cfg, "hn_n_layers", 3))
    heads = _get(cfg, "hn_num_attention_heads", None) or H // 64
    separate = bool(_get(cfg, "separate_out_embeddings", False))
    eps = 1e-5
    # modeling_hypernet.py:170-188
    use_fb = ids >= v0
    x = F.embedding(ids.clamp(max=v0 - 1), source_embeddings)
    if _get(cfg, "hn_rescale_embeddings", False):
        x = W["in_scaler.w"] * x + W["in_scaler.b"]
    x = torch.where(use_fb[..., None], F.embedding((ids - v0).clamp(min=0), W["fallback_embeddings.weight"]), x)
    x = F.linear(x, W["input_projection.0.weight"], W["input_projection.0.bias"])  # :189
    x = _projector(x, W, "input_projection.1.")
    mask = ids != pad
    if _get(cfg, "hn_embed_lang_id", False):  # :192-218
        L = mask.shape[1]
        lang = W["lang_embeddings.weight"][int(lang_index)] - (
            W["model.embeddings.token_type_embeddings.weight"][0] + W["model.embeddings.position_embeddings.weight"][L])
        x = torch.cat([x, lang[None, None, :].expand(x.shape[0], -1, -1)], dim=1)
        mask = torch.cat([mask, torch.ones((mask.shape[0], 1), dtype=torch.bool, device=dev)], dim=1)
    B, S, _ = x.shape
    # RobertaEmbeddings + encoder (eager attention)
    x = x + W["model.embeddings.token_type_embeddings.weight"][0]
    x = x + W["model.embeddings.position_embeddings.weight"][:S][None]
    x = F.layer_norm(x, (H,), W["model.embeddings.LayerNorm.weight"], W["model.embeddings.LayerNorm.bias"], eps)
    add_mask = torch.zeros((B, 1, 1, S), dtype=x.dtype, device=dev).masked_fill(~mask[:, None, None, :], torch.finfo(x.dtype).min)
    dh = H // heads
    for l in range(n_layers):
        p = f"model.encoder.layer.{l}."
        q = F.linear(x, W[p + "attention.self.query.weight"], W[p + "attention.self.query.bias"]).view(B, S, heads, dh).transpose(1, 2)
        k = F.linear(x, W[p + "attention.self.key.weight"], W[p + "attention.self.key.bias"]).view(B, S, heads, dh).transpose(1, 2)
        v = F.linear(x, W[p + "attention.self.value.weight"], W[p + "attention.self.value.bias"]).view(B, S, heads, dh).transpose(1, 2)
        s = torch.matmul(q, k.transpose(2, 3)) * (dh ** -0.5) + add_mask
        ctx = torch.matmul(F.softmax(s, dim=-1), v).transpose(1, 2).reshape(B, S, H)
        a = F.linear(ctx, W[p + "attention.output.dense.weight"], W[p + "attention.output.dense.bias"])
        x = F.layer_norm(a + x, (H,), W[p + "attention.output.LayerNorm.weight"], W[p + "attention.output.LayerNorm.bias"], eps)
        h = F.gelu(F.linear(x, W[p + "intermediate.dense.weight"], W[p + "intermediate.dense.bias"]))
        h = F.linear(h, W[p + "output.dense.weight"], W[p + "output.dense.bias"])
        x = F.layer_norm(h + x, (H,), W[p + "output.LayerNorm.weight"], W[p + "output.LayerNorm.bias"], eps)
    h0 = x[:, 0]  # :231-234

    def head(prefix):
        return F.linear(_projector(h0, W, prefix + "0."), W[prefix + "1.weight"], W[prefix + "1.bias"])

    pred = head("output_projection.")
    if _get(cfg, "hn_single_head", False):
        pred_in, pred_out = pred[..., :D], (pred[..., D:] if separate else None)
    else:
        pred_in, pred_out = pred, (head("output_projection_out.") if separate else None)
    if _get(cfg, "hn_rescale_embeddings", False):
        pred_in = W["scaler.w"] * pred_in + W["scaler.b"]
        if pred_out is not None:
            pred_out = W["out_scaler.w"] * pred_out + W["out_scaler.b"]
    if _get(cfg, "hn_predict_bias", False):
        bias = F.linear(h0, W["bias_projection.weight"], W["bias_projection.bias"])[..., 0]
    else:
        bias = torch.zeros(ids.shape[0], device=dev)
    if dev.type != "cpu":  # library-kernel comparison arm of bench.py: results stay on the device
        return pred_in, pred_out, bias
    return pred_in.numpy(), (None if pred_out is None else pred_out.numpy()), bias.numpy()
