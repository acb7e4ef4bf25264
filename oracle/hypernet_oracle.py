"""CPU oracle for the ZeTT hypernetwork forward  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product path (``zett_b200``) never does; it fails loudly when the
CUDA library is missing.

This is a numpy restatement of the reference's PyTorch binding of the hypernetwork,
``hf_hypernet/modeling_hypernet.py:156-267`` (``ZettHypernet.__call__``), including the parts of
``transformers.RobertaModel`` (third-party, transformers==4.45.2 pinned by ``requirements.txt:2``;
5.5.0 installed) that the reference reaches through ``self.model(inputs_embeds=..., attention_mask=...,
position_ids=...)`` at ``modeling_hypernet.py:225-229``.

Parity pin: the reference holds no golden vectors for this path.  This oracle is
pinned against outputs of the reference itself, generated in the build container by
``tests/golden/make_golden.py`` (imports ``/root/reference/hf_hypernet``) and committed as
``tests/golden/hypernet_*.npz``; ``tests/test_oracle_float.py`` checks oracle == reference on them.

Weights are passed as a flat ``{state_dict name: ndarray}`` mapping using exactly the reference's
``state_dict`` names (weight-name contract).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np

try:  # scipy ships in the image; keep a pure-numpy erf for safety
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf, otypes=[np.float64])


def _cfg_get(cfg, name, default=None):
    if isinstance(cfg, dict):
        return cfg.get(name, default)
    return getattr(cfg, name, default)


def gelu_tanh(x):
    """``F.gelu(x, approximate="tanh")`` -- ProjectorBlock, modeling_hypernet.py:36-39."""
    c = x.dtype.type(math.sqrt(2.0 / math.pi))
    return x.dtype.type(0.5) * x * (x.dtype.type(1.0) + np.tanh(c * (x + x.dtype.type(0.044715) * x * x * x)))


def gelu_erf(x):
    """Exact GELU (``hidden_act="gelu"``) -- RobertaIntermediate."""
    return (x.dtype.type(0.5) * x * (x.dtype.type(1.0) + _erf(x / x.dtype.type(math.sqrt(2.0))))).astype(x.dtype)


def layer_norm(x, w, b, eps):
    mu = x.mean(axis=-1, keepdims=True)
    xc = x - mu
    var = (xc * xc).mean(axis=-1, keepdims=True)
    return xc / np.sqrt(var + x.dtype.type(eps)) * w + b


def linear(x, w, b):
    """``nn.Linear``: y = x W^T + b with W [out, in] (leading dims flattened into one GEMM, as ATen does)."""
    y = x.reshape(-1, x.shape[-1]) @ w.T
    return y.reshape(x.shape[:-1] + (w.shape[0],)) + b


def projector_block(x, W, prefix):
    """``ProjectorBlock.__call__`` -- modeling_hypernet.py:35-40 (both GELUs tanh, LN eps 1e-6)."""
    h = gelu_tanh(linear(x, W[prefix + "dense1.weight"], W[prefix + "dense1.bias"]))
    h = gelu_tanh(linear(h, W[prefix + "dense2.weight"], W[prefix + "dense2.bias"]))
    return layer_norm(h + x, W[prefix + "ln.weight"], W[prefix + "ln.bias"], 1e-6)


def roberta_encoder(x, key_mask, W, n_layers, n_heads, ln_eps=1e-5):
    """RobertaModel(add_pooling_layer=False) on ``inputs_embeds`` with explicit arange position ids.

    x: [B, S, H] inputs_embeds; key_mask: [B, S] bool.  Eager attention semantics: additive
    ``finfo.min`` on masked keys (a fully masked row therefore attends uniformly over all S keys).
    """
    B, S, H = x.shape
    dt = x.dtype
    dh = H // n_heads
    # RobertaEmbeddings: inputs_embeds + token_type_embeddings[0] + position_embeddings[pos]; LayerNorm
    x = x + W["model.embeddings.token_type_embeddings.weight"][0]
    x = x + W["model.embeddings.position_embeddings.weight"][:S][None]
    x = layer_norm(x, W["model.embeddings.LayerNorm.weight"], W["model.embeddings.LayerNorm.bias"], ln_eps)
    add_mask = np.where(key_mask, dt.type(0.0), np.finfo(dt).min).astype(dt)[:, None, None, :]
    scaling = dt.type(dh ** -0.5)
    for l in range(n_layers):
        p = f"model.encoder.layer.{l}."
        q = linear(x, W[p + "attention.self.query.weight"], W[p + "attention.self.query.bias"])
        k = linear(x, W[p + "attention.self.key.weight"], W[p + "attention.self.key.bias"])
        v = linear(x, W[p + "attention.self.value.weight"], W[p + "attention.self.value.bias"])
        q = q.reshape(B, S, n_heads, dh).transpose(0, 2, 1, 3)
        k = k.reshape(B, S, n_heads, dh).transpose(0, 2, 1, 3)
        v = v.reshape(B, S, n_heads, dh).transpose(0, 2, 1, 3)
        s = (q @ k.transpose(0, 1, 3, 2)) * scaling + add_mask
        s = s - s.max(axis=-1, keepdims=True)
        e = np.exp(s)
        pr = e / e.sum(axis=-1, keepdims=True)
        ctx = (pr @ v).transpose(0, 2, 1, 3).reshape(B, S, H)
        a = linear(ctx, W[p + "attention.output.dense.weight"], W[p + "attention.output.dense.bias"])
        x = layer_norm(a + x, W[p + "attention.output.LayerNorm.weight"], W[p + "attention.output.LayerNorm.bias"], ln_eps)
        h = gelu_erf(linear(x, W[p + "intermediate.dense.weight"], W[p + "intermediate.dense.bias"]))
        h = linear(h, W[p + "output.dense.weight"], W[p + "output.dense.bias"])
        x = layer_norm(h + x, W[p + "output.LayerNorm.weight"], W[p + "output.LayerNorm.bias"], ln_eps)
    return x


def hypernet_forward(
    cfg,
    weights: Dict[str, np.ndarray],
    target_surface_forms: np.ndarray,
    source_embeddings: np.ndarray,
    lang_index: Optional[int] = None,
    dtype=np.float32,
) -> Tuple[np.ndarray, Optional[np.ndarray], np.ndarray]:
    """Restatement of ``ZettHypernet.__call__`` (modeling_hypernet.py:156-267).

    Returns ``(pred_in [V, D], pred_out [V, D] | None, pred_bias [V])`` in ``dtype``.
    """
    dt = np.dtype(dtype)
    W = {k: np.asarray(v, dtype=dt) for k, v in weights.items()}
    ids = np.asarray(target_surface_forms).astype(np.int64)
    src = np.asarray(source_embeddings, dtype=dt)

    if _cfg_get(cfg, "hn_model_type", "roberta") != "roberta":
        raise NotImplementedError()  # :78-79
    if _cfg_get(cfg, "hn_add_inter_token_attention", False) or _cfg_get(cfg, "hn_embed_target_priors", False):
        raise NotImplementedError()  # :85-89
    if not _cfg_get(cfg, "hn_embed_using_source_embeddings", False):
        raise NotImplementedError()  # :167-168

    v0 = int(_cfg_get(cfg, "original_vocab_size"))
    pad = int(_cfg_get(cfg, "pad_token_id"))
    H = int(_cfg_get(cfg, "hn_hidden_size"))
    D = int(_cfg_get(cfg, "n_embd"))
    n_layers = int(_cfg_get(cfg, "hn_n_layers", 3))
    n_heads = _cfg_get(cfg, "hn_num_attention_heads", None) or H // 64  # :73-75
    separate = bool(_cfg_get(cfg, "separate_out_embeddings", False))

    # :170-188  clamp / fallback split, gather, rescale, select
    use_fb = ids >= v0
    main_ids = np.minimum(ids, v0 - 1)
    fb_ids = np.maximum(ids - v0, 0)
    x = src[main_ids]
    if _cfg_get(cfg, "hn_rescale_embeddings", False):
        x = W["in_scaler.w"] * x + W["in_scaler.b"]
    x = np.where(use_fb[..., None], W["fallback_embeddings.weight"][fb_ids], x)

    # :189  input_projection = Linear(E, H) ; ProjectorBlock(H, H, I)
    x = linear(x, W["input_projection.0.weight"], W["input_projection.0.bias"])
    x = projector_block(x, W, "input_projection.1.")
    mask = ids != pad  # :190

    if _cfg_get(cfg, "hn_embed_lang_id", False):  # :192-218
        L = mask.shape[1]
        lang = W["lang_embeddings.weight"][int(lang_index)].copy()
        lang = lang - (
            W["model.embeddings.token_type_embeddings.weight"][0]
            + W["model.embeddings.position_embeddings.weight"][L]
        )
        x = np.concatenate([x, np.broadcast_to(lang[None, None, :], (x.shape[0], 1, H))], axis=1)
        mask = np.concatenate([mask, np.ones((mask.shape[0], 1), dtype=bool)], axis=1)

    hidden = roberta_encoder(x, mask, W, n_layers, n_heads)  # :225-229

    if _cfg_get(cfg, "hn_concat_last_hidden_state", False):
        raise NotImplementedError("shape-inconsistent in the reference PyTorch module (SURVEY 8a row a9)")
    h0 = hidden[:, 0]  # :231-234

    def head(prefix):
        y = projector_block(h0, W, prefix + "0.")
        return linear(y, W[prefix + "1.weight"], W[prefix + "1.bias"])

    pred = head("output_projection.")  # :236
    if _cfg_get(cfg, "hn_single_head", False):  # :238-246
        pred_in = pred[..., :D]
        pred_out = pred[..., D:] if separate else None
    else:
        pred_in = pred
        pred_out = head("output_projection_out.") if separate else None

    if _cfg_get(cfg, "hn_rescale_embeddings", False):  # :254-258
        pred_in = W["scaler.w"] * pred_in + W["scaler.b"]
        if pred_out is not None:
            pred_out = W["out_scaler.w"] * pred_out + W["out_scaler.b"]

    if _cfg_get(cfg, "hn_predict_bias", False):  # :260-265
        bias = linear(h0, W["bias_projection.weight"], W["bias_projection.bias"])[..., 0]
    else:
        bias = np.zeros(ids.shape[0], dtype=dt)
    return pred_in.astype(dt), (None if pred_out is None else pred_out.astype(dt)), bias.astype(dt)


def fully_masked_rows(cfg, target_surface_forms) -> np.ndarray:
    """Rows whose every key is masked (all ids == pad and no lang-id slot).  The reference's result for
    those rows depends on the attention backend (SURVEY 8a edge semantics); they are reported separately."""
    ids = np.asarray(target_surface_forms)
    if _cfg_get(cfg, "hn_embed_lang_id", False):
        return np.zeros(ids.shape[0], dtype=bool)
    return (ids == int(_cfg_get(cfg, "pad_token_id"))).all(axis=1)


def rel_errors(x: np.ndarray, ref: np.ndarray, exclude: Optional[np.ndarray] = None):
    """(Frobenius-relative error, worst-row relative error) as defined in SURVEY 8d 'Parity'."""
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if exclude is not None and exclude.any():
        x, ref = x[~exclude], ref[~exclude]
    fro = float(np.linalg.norm(x - ref) / max(np.linalg.norm(ref), 1e-30))
    if x.ndim == 1:
        return fro, fro
    rn = np.linalg.norm(ref, axis=1)
    worst = float((np.linalg.norm(x - ref, axis=1) / np.maximum(rn, 1e-30)).max()) if len(rn) else 0.0
    return fro, worst
