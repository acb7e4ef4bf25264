"""``ZettHypernet`` -- the reference's call surface (hf_hypernet/modeling_hypernet.py:43-267) over the B200 kernels.

    hypernet = AutoModel.from_pretrained(path, trust_remote_code=True)            # or ZettHypernet(config)
    pred_in, pred_out, pred_bias = hypernet(target_surface_forms, source_embeddings=src, lang_index=lang)

Same constructor argument (a ``ZettHypernetConfig``), same ``state_dict`` names and shapes (so the reference's
checkpoints load unchanged), same call signature, outputs and exception types.  The torch side is I/O only: the
parameters are plain containers that are handed once to ``libzett_b200.so`` (``zett_hn_set_weight`` /
``zett_hn_finalize``); the forward is one ``zett_hn_forward`` call on the current CUDA stream -- no ``torch.nn`` op
runs on the hot path and there is no CPU fallback (a CPU-only process raises).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Tuple

import numpy as np
import torch
from torch import nn
from transformers import PreTrainedModel

from . import _lib
from .config import ROBERTA_LAYER_NORM_EPS, ROBERTA_MAX_POSITION_EMBEDDINGS, ZettHypernetConfig, weight_shapes

class _Holder(nn.Module):
    """A node of the parameter tree; carries no computation."""


def make_c_config(cfg: ZettHypernetConfig, max_rows_per_pass: int = 0, gemm_impl: int = 0,
                  split_terms: int = 0) -> _lib.ZettHnConfig:
    c = _lib.ZettHnConfig()
    c.struct_bytes = ctypes.sizeof(_lib.ZettHnConfig)
    c.hn_surface_maxlen = int(cfg.hn_surface_maxlen)
    c.hn_n_layers = int(cfg.hn_n_layers)
    c.n_embd = int(cfg.n_embd)
    c.hn_hidden_size = int(cfg.hn_hidden_size or 0)
    c.hn_intermediate_size = int(cfg.hn_intermediate_size or 0)
    c.hn_num_attention_heads = int(cfg.hn_num_attention_heads or 0)
    c.hn_rescale_embeddings = int(bool(cfg.hn_rescale_embeddings))
    c.hn_embed_target_priors = int(bool(cfg.hn_embed_target_priors))
    c.hn_add_inter_token_attention = int(bool(cfg.hn_add_inter_token_attention))
    c.hn_embed_using_source_embeddings = int(bool(cfg.hn_embed_using_source_embeddings))
    c.hn_concat_last_hidden_state = int(bool(cfg.hn_concat_last_hidden_state))
    c.hn_single_head = int(bool(cfg.hn_single_head))
    c.hn_predict_bias = int(bool(getattr(cfg, "hn_predict_bias", False)))
    c.hn_embed_lang_id = int(bool(cfg.hn_embed_lang_id))
    c.hn_model_type_is_roberta = int(cfg.hn_model_type == "roberta")
    c.n_langs = int(cfg.n_langs or 0)
    c.pad_token_id = int(cfg.pad_token_id)
    c.original_vocab_size = int(cfg.original_vocab_size or 0)
    c.hn_n_extra_tokens = int(cfg.hn_n_extra_tokens or 0)
    c.separate_out_embeddings = int(bool(cfg.separate_out_embeddings))
    c.max_position_embeddings = ROBERTA_MAX_POSITION_EMBEDDINGS
    c.encoder_layer_norm_eps = ROBERTA_LAYER_NORM_EPS
    c.max_rows_per_pass = int(max_rows_per_pass)
    c.gemm_impl = int(gemm_impl)
    c.split_terms = int(split_terms)
    return c


class NativeHypernet:
    """Owner of one ``zett_hn`` handle (one CUDA device).  Usable without the HF wrapper: weights come as a
    ``{state_dict name: tensor or ndarray}`` mapping."""

    def __init__(self, cfg: ZettHypernetConfig, weights, device: torch.device, max_rows_per_pass: int = 0,
                 gemm_impl: int = 0, split_terms: int = 0):
        if device.type != "cuda":
            raise RuntimeError("zett_b200 runs on a B200 GPU only; there is no CPU fallback (got device %s)" % device)
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = device
        self.auto_terms = int(split_terms) == 0   # the caller left the operand format to the library: fallback allowed
        self.handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(self.lib.zett_hn_create(ctypes.byref(make_c_config(cfg, max_rows_per_pass, gemm_impl, split_terms)),
                                               ctypes.byref(self.handle)))
            try:
                for name, w in weights.items():
                    if name.endswith("position_ids") or name.endswith("token_type_ids"):
                        continue
                    t = torch.as_tensor(w) if not isinstance(w, torch.Tensor) else w.detach()
                    code = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}.get(t.dtype)
                    if code is None:
                        t, code = t.float(), _lib.F32
                    t = t.contiguous()
                    shape = (ctypes.c_int64 * max(t.dim(), 1))(*t.shape)
                    _lib.check(self.lib.zett_hn_set_weight(self.handle, name.encode(), ctypes.c_void_p(t.data_ptr()), code,
                                                           t.dim(), shape))
                _lib.check(self.lib.zett_hn_finalize(self.handle))
            except Exception:
                self.close()
                raise

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.zett_hn_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def forward_into(self, surface_forms: torch.Tensor, source_embeddings: torch.Tensor, lang_index: int,
                     pred_in: torch.Tensor, pred_out: Optional[torch.Tensor], pred_bias: torch.Tensor,
                     ld_pred: int = 0, ld_bias: int = 0):
        """Raw launch on the current stream; all tensors on ``self.device``; no synchronisation."""
        n = surface_forms.shape[0]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _lib.check(self.lib.zett_hn_forward(
                self.handle, ctypes.c_void_p(surface_forms.data_ptr()), n, ctypes.c_void_p(source_embeddings.data_ptr()),
                source_embeddings.shape[0], int(lang_index), ctypes.c_void_p(pred_in.data_ptr()),
                ctypes.c_void_p(pred_out.data_ptr() if pred_out is not None else 0), ctypes.c_void_p(pred_bias.data_ptr()),
                int(ld_pred), int(ld_bias), ctypes.c_void_p(stream)))

    def check(self):
        """Synchronise the current stream and raise what the kernels recorded (IndexError for out-of-range ids)."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.zett_hn_check(self.handle, ctypes.c_void_p(stream)))

    def set_split_terms(self, terms: int):
        """Switch the operand format of the handle (``zett_hn_set_split_terms``); synchronises the device."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.zett_hn_set_split_terms(self.handle, int(terms)))

    def run_checked(self, enqueue, allow_fallback: bool = True):
        """``enqueue()`` (any number of ``forward_into`` calls) followed by ``check()``.  When the kernels report an operand
        outside fp16's range (``OperandRangeError``) and the operand format was left on auto, the handle switches to the
        three-term bf16 split -- fp32's exponent range, the semantics of the fp32 reference
        (hf_hypernet/modeling_hypernet.py:179-189) -- and the work is enqueued and checked again."""
        enqueue()
        try:
            self.check()
        except _lib.OperandRangeError as e:
            if not (self.auto_terms and allow_fallback):
                raise
            import warnings
            warnings.warn("zett_b200: %s -- switching this model to split_terms = 3 (three bf16 terms) and repeating the forward" % e)
            self.set_split_terms(3)
            enqueue()
            self.check()

    def stats(self) -> dict:
        st = _lib.ZettHnStats()
        _lib.check(self.lib.zett_hn_get_stats(self.handle, ctypes.byref(st)))
        return {k: getattr(st, k) for k, _ in st._fields_}

    def set_timing(self, enable: bool):
        _lib.check(self.lib.zett_hn_set_timing(self.handle, int(bool(enable))))

    def workspace_bytes(self, n_rows: int) -> int:
        return int(self.lib.zett_hn_workspace_bytes(self.handle, n_rows))


class ZettHypernet(PreTrainedModel):
    """Drop-in for the reference's ``hf_hypernet.ZettHypernet`` (same config, weights, call and return values)."""

    config_class = ZettHypernetConfig
    base_model_prefix = ""
    main_input_name = "target_surface_forms"
    _supports_sdpa = False

    # execution knobs (not part of the reference's surface)
    max_rows_per_pass = 0      # 0 -> 16384 rows per pass of the kernels (the reference's transfer batch size)
    gemm_impl = 0              # see zett_hn_config.gemm_impl
    split_terms = 0            # see zett_hn_config.split_terms
    check_ids = True           # synchronise after each call and raise IndexError for out-of-range ids

    def __init__(self, config: ZettHypernetConfig):
        super().__init__(config)
        self.config = config
        # the branches the reference rejects at construction time (modeling_hypernet.py:78-79, 85-89)
        if config.hn_model_type != "roberta":
            raise NotImplementedError()
        if config.hn_add_inter_token_attention or config.hn_embed_target_priors:
            raise NotImplementedError()
        if config.hn_concat_last_hidden_state:
            raise NotImplementedError("hn_concat_last_hidden_state is shape-inconsistent in the reference PyTorch module")
        self.has_separate_out_embeddings = getattr(config, "separate_out_embeddings", False)
        if getattr(config, "hn_num_attention_heads", None) is None:
            config.hn_num_attention_heads = config.hn_hidden_size // 64  # modeling_hypernet.py:73-75
        self.pad_token_id = config.pad_token_id
        assert self.pad_token_id is not None  # modeling_hypernet.py:92
        for name, shape in weight_shapes(config).items():
            self._add_parameter(name, shape)
        self._native: Optional[NativeHypernet] = None
        self.post_init()

    # ---- parameter containers -----------------------------------------------------------------------------------
    def _add_parameter(self, dotted: str, shape):
        parts = dotted.split(".")
        mod = self
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, _Holder())
            mod = mod._modules[p]
        mod.register_parameter(parts[-1], nn.Parameter(torch.zeros(shape, dtype=torch.float32), requires_grad=False))

    def _init_weights(self, module):  # parameters come from a checkpoint; nothing to initialise
        return

    def _apply(self, fn, *args, **kwargs):  # .to() / .cuda() / .float(): the native copy is rebuilt lazily
        self._drop_native()
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._drop_native()
        return super().load_state_dict(*args, **kwargs)

    def _drop_native(self):
        nat = self.__dict__.get("_native")
        if nat is not None:
            nat.close()
            self._native = None

    def refresh(self):
        """Call after mutating parameters in place; the next forward re-uploads them to the kernels."""
        self._drop_native()

    def native(self, device: Optional[torch.device] = None) -> NativeHypernet:
        if device is None:
            device = next(self.parameters()).device
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError(
                "zett_b200.ZettHypernet computes on a B200 GPU only (no CPU fallback): move the model or its inputs to "
                "a CUDA device")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if self._native is None or self._native.device != device:
            self._drop_native()
            self._native = NativeHypernet(self.config, self.state_dict(), device, self.max_rows_per_pass, self.gemm_impl,
                                          self.split_terms)
        return self._native

    # ---- the reference's call ------------------------------------------------------------------------------------
    def forward(self, target_surface_forms, target_priors=None, source_embeddings=None, lang_index=None,
                deterministic: bool = True):
        if target_priors is not None:
            raise NotImplementedError()  # modeling_hypernet.py:164-165
        if not self.config.hn_embed_using_source_embeddings:
            raise NotImplementedError()  # modeling_hypernet.py:167-168
        if source_embeddings is None:
            raise ValueError("source_embeddings is required")
        param_device = next(self.parameters()).device
        if param_device.type == "cuda":
            device = param_device
        elif isinstance(source_embeddings, torch.Tensor) and source_embeddings.is_cuda:
            device = source_embeddings.device
        elif isinstance(target_surface_forms, torch.Tensor) and target_surface_forms.is_cuda:
            device = target_surface_forms.device
        else:
            device = param_device  # -> native() raises: no CPU path
        nat = self.native(device)
        device = nat.device
        sf = torch.as_tensor(target_surface_forms)
        squeeze = sf.dim() == 1
        if squeeze:
            sf = sf[None]
        if sf.dim() != 2 or sf.shape[1] != self.config.hn_surface_maxlen:
            raise ValueError("target_surface_forms must be [n, hn_surface_maxlen=%d]" % self.config.hn_surface_maxlen)
        sf = sf.to(device=device, dtype=torch.int32, non_blocking=True).contiguous()
        src = torch.as_tensor(source_embeddings).to(device=device, dtype=torch.float32, non_blocking=True).contiguous()
        if src.dim() != 2 or src.shape[1] != self.config.n_in_embd:
            raise ValueError("source_embeddings must be [rows, %d]" % self.config.n_in_embd)
        if self.config.hn_embed_lang_id:
            if lang_index is None:
                raise ValueError("lang_index is required when hn_embed_lang_id is set")
            lang = int(lang_index.item()) if isinstance(lang_index, torch.Tensor) else int(lang_index)
        else:
            lang = -1
        n, D = sf.shape[0], self.config.n_embd
        pred_in = torch.empty((n, D), dtype=torch.float32, device=device)
        pred_out = torch.empty((n, D), dtype=torch.float32, device=device) if self.has_separate_out_embeddings else None
        pred_bias = torch.empty((n,), dtype=torch.float32, device=device)
        if n > 0:
            if self.check_ids:
                nat.run_checked(lambda: nat.forward_into(sf, src, lang, pred_in, pred_out, pred_bias))
            else:
                nat.forward_into(sf, src, lang, pred_in, pred_out, pred_bias)
        if squeeze:
            return pred_in[0], (None if pred_out is None else pred_out[0]), pred_bias[0]
        return pred_in, pred_out, pred_bias

    __call__ = nn.Module.__call__


def load_weights_numpy(model: ZettHypernet, weights: Dict[str, np.ndarray]):
    """Fill the parameter containers from a ``{name: ndarray}`` mapping (tests / benchmarks)."""
    sd = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in weights.items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    if unexpected or missing:
        raise ValueError("state_dict mismatch: missing=%s unexpected=%s" % (missing, unexpected))
    return model


def register_auto_classes():
    """``AutoConfig`` / ``AutoModel`` resolve ``model_type == "zett_hypernetwork"`` to the B200 classes, so
    ``AutoModel.from_pretrained(path)`` (README.md:93-117 of the reference) returns a ``zett_b200.ZettHypernet``.
    With ``trust_remote_code=True`` and an ``auto_map`` in the checkpoint HF prefers the checkpoint's own code; pass
    ``trust_remote_code=False`` (or use ``ZettHypernet.from_pretrained``) to take this implementation."""
    from transformers import AutoConfig, AutoModel
    AutoConfig.register(ZettHypernetConfig.model_type, ZettHypernetConfig, exist_ok=True)
    AutoModel.register(ZettHypernetConfig, ZettHypernet, exist_ok=True)
