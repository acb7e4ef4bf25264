"""Import the reference's own modules staged by ``oracle/make_ref.py`` -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

Two stubs make them importable offline: ``RobertaConfig.from_pretrained("roberta-base")`` (a hub
fetch, hf_hypernet/modeling_hypernet.py:67-69) returns the roberta-base constants, and the jax / flax / optax modules
that ``zett/utils.py`` imports at module top (zett/utils.py:3-20, 26) but ``get_surface_form_matrix`` never touches are
replaced by mocks.
"""
import os
import sys
from unittest.mock import MagicMock

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "hf_hypernet", "modeling_hypernet.py"))


def _path():
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)


def load_hypernet():
    """(ZettHypernetConfig, ZettHypernet) of the reference (hf_hypernet/)."""
    from transformers import RobertaConfig
    RobertaConfig.from_pretrained = classmethod(lambda cls, *a, **k: RobertaConfig(
        vocab_size=50265, max_position_embeddings=514, type_vocab_size=1, layer_norm_eps=1e-5,
        pad_token_id=1, bos_token_id=0, eos_token_id=2))
    _path()
    from hf_hypernet.configuration_hypernet import ZettHypernetConfig as RefConfig
    from hf_hypernet.modeling_hypernet import ZettHypernet as RefHypernet
    return RefConfig, RefHypernet


def load_utils():
    """The reference's ``zett.utils`` (get_surface_form_matrix, CHARS_TO_BYTES)."""
    for n in ["jax", "jax.numpy", "jax.sharding", "flax", "flax.linen", "flax.serialization", "flax.traverse_util", "optax"]:
        sys.modules.setdefault(n, MagicMock())
    sys.modules["flax.linen"].Module = type("Module", (), {})
    _path()
    import zett.utils as ref_utils
    return ref_utils


def build_hypernet(cfg, weights, device="cpu"):
    """The reference module with the given config (a zett_b200 ZettHypernetConfig) and ``{name: ndarray}`` weights,
    eager attention (deterministic masked-row semantics)."""
    import torch
    RefConfig, RefHypernet = load_hypernet()
    ref_cfg = RefConfig(**{k: v for k, v in cfg.to_dict().items()
                           if k.startswith("hn_") or k in ("n_embd", "n_langs", "pad_token_id", "original_vocab_size",
                                                           "separate_out_embeddings", "use_unigram_bias")})
    model = RefHypernet(ref_cfg).eval()
    model.model.config._attn_implementation = "eager"
    missing, unexpected = model.load_state_dict({k: torch.as_tensor(v) for k, v in weights.items()}, strict=False)
    assert not unexpected, unexpected
    assert all("position_ids" in m or "token_type_ids" in m or "word_embeddings" in m for m in missing), missing
    return model.to(device)
