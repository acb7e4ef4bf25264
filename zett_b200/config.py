"""``ZettHypernetConfig`` -- same field names and defaults as the reference's config
(hf_hypernet/configuration_hypernet.py:3-56) so a trained checkpoint's ``config.json`` loads unchanged.

Fields that training writes onto the config without declaring them (SURVEY.md section 8b, config contract):
``original_vocab_size`` (train.py:314), ``hn_n_extra_tokens`` (train.py:361), ``separate_out_embeddings``
(train.py:350), ``pad_token_id`` (train.py:295), ``langs`` (scripts/convert_to_pt.py:31-33) ride in through
``**kwargs`` exactly as they do in the reference.
"""
from __future__ import annotations

from typing import Dict, Tuple

from transformers import PretrainedConfig

# roberta-base constants that the reference pulls from the hub (hf_hypernet/modeling_hypernet.py:67-69)
ROBERTA_MAX_POSITION_EMBEDDINGS = 514
ROBERTA_LAYER_NORM_EPS = 1e-5

# (field, default) in the reference's declaration order
_HN_FIELDS = (
    ("hn_model_name_or_path", "roberta-base"),
    ("hn_surface_maxlen", 16),
    ("hn_n_layers", 3),
    ("n_embd", 768),
    ("hn_hidden_size", None),
    ("hn_intermediate_size", None),
    ("hn_rescale_embeddings", False),
    ("use_unigram_bias", False),
    ("hn_embed_target_priors", False),
    ("hn_add_inter_token_attention", False),
    ("hn_inter_token_attention_bias_by_priors", False),
    ("hn_inter_token_attention_bias_scaler", 1.0),
    ("hn_n_inter_token_blocks", 16),
    ("hn_language_adapter_bottleneck_dim", 0),
    ("hn_embed_using_source_embeddings", False),
    ("hn_concat_last_hidden_state", False),
    ("hn_single_head", False),
    ("hn_predict_bias", True),
    ("hn_num_attention_heads", None),
    ("hn_embed_lang_id", False),
    ("hn_model_type", "roberta"),
    ("n_langs", None),
)

# set by the reference's training script, not constructor arguments there either
_EXTRA_FIELDS = (
    ("original_vocab_size", None),
    ("hn_n_extra_tokens", 0),
    ("separate_out_embeddings", False),
)


class ZettHypernetConfig(PretrainedConfig):
    model_type = "zett_hypernetwork"

    def __init__(self, **kwargs):
        values = {}
        for name, default in _HN_FIELDS + _EXTRA_FIELDS:
            values[name] = kwargs.pop(name, default)
        super().__init__(**kwargs)
        for name, value in values.items():
            setattr(self, name, value)
        self.model_type = "zett_hypernetwork"

    # derived quantities used throughout the B200 path ------------------------------------------------
    @property
    def n_in_embd(self) -> int:
        """E: width of a source-embedding row (hf_hypernet/modeling_hypernet.py:59-64)."""
        return self.n_embd * 2 if self.separate_out_embeddings else self.n_embd

    @property
    def n_heads(self) -> int:
        """hf_hypernet/modeling_hypernet.py:73-75."""
        return self.hn_num_attention_heads or self.hn_hidden_size // 64

    @property
    def seq_len(self) -> int:
        """S = L (+1 when a lang-id slot is appended, modeling_hypernet.py:192-218)."""
        return self.hn_surface_maxlen + (1 if self.hn_embed_lang_id else 0)

    def flops_per_row(self, lengths=None, pruned: bool = True) -> float:
        """Algorithmic FLOPs per vocabulary row.

        ``lengths=None``: dense F_ref (``pruned=False``) or F_min (last layer pruned to row 0).
        ``lengths=l``: executed work when only ``l`` of the L surface positions (+ lang slot) are active.
        """
        H, I, D, E = self.hn_hidden_size, self.hn_intermediate_size, self.n_embd, self.n_in_embd
        L = self.hn_surface_maxlen if lengths is None else lengths
        S = L + (1 if self.hn_embed_lang_id else 0)
        n_layers = self.hn_n_layers
        heads_out = 2 if (self.separate_out_embeddings and not self.hn_single_head) else 1
        d_out = E if self.hn_single_head else D
        f = L * (2 * E * H + 4 * H * I)
        full_layer = S * (8 * H * H + 4 * H * I + 4 * S * H)
        if pruned:
            last = S * 4 * H * H + (4 * H * H + 4 * H * I) + 4 * S * H
            f += (n_layers - 1) * full_layer + last
        else:
            f += n_layers * full_layer
        f += heads_out * (4 * H * I + 2 * H * d_out) + (2 * H if self.hn_predict_bias else 0)
        return float(f)


def weight_shapes(cfg: "ZettHypernetConfig") -> Dict[str, Tuple[int, ...]]:
    """The reference's ``state_dict`` names and shapes (hf_hypernet/modeling_hypernet.py:46-154)."""
    H, I, D = cfg.hn_hidden_size, cfg.hn_intermediate_size, cfg.n_embd
    E = cfg.n_in_embd
    s: Dict[str, Tuple[int, ...]] = {}
    s["model.embeddings.word_embeddings.weight"] = (cfg.pad_token_id + 1, H)
    s["model.embeddings.token_type_embeddings.weight"] = (1, H)
    s["model.embeddings.position_embeddings.weight"] = (ROBERTA_MAX_POSITION_EMBEDDINGS, H)
    s["model.embeddings.LayerNorm.weight"] = (H,)
    s["model.embeddings.LayerNorm.bias"] = (H,)
    for l in range(cfg.hn_n_layers):
        p = f"model.encoder.layer.{l}."
        for n in ("query", "key", "value"):
            s[p + f"attention.self.{n}.weight"] = (H, H)
            s[p + f"attention.self.{n}.bias"] = (H,)
        s[p + "attention.output.dense.weight"] = (H, H)
        s[p + "attention.output.dense.bias"] = (H,)
        s[p + "attention.output.LayerNorm.weight"] = (H,)
        s[p + "attention.output.LayerNorm.bias"] = (H,)
        s[p + "intermediate.dense.weight"] = (I, H)
        s[p + "intermediate.dense.bias"] = (I,)
        s[p + "output.dense.weight"] = (H, I)
        s[p + "output.dense.bias"] = (H,)
        s[p + "output.LayerNorm.weight"] = (H,)
        s[p + "output.LayerNorm.bias"] = (H,)
    s["fallback_embeddings.weight"] = (max(cfg.hn_n_extra_tokens, 1), E)
    s["input_projection.0.weight"] = (H, E)
    s["input_projection.0.bias"] = (H,)

    def projector(prefix):
        s[prefix + "dense1.weight"] = (I, H)
        s[prefix + "dense1.bias"] = (I,)
        s[prefix + "dense2.weight"] = (H, I)
        s[prefix + "dense2.bias"] = (H,)
        s[prefix + "ln.weight"] = (H,)
        s[prefix + "ln.bias"] = (H,)

    projector("input_projection.1.")
    projector("output_projection.0.")
    s["output_projection.1.weight"] = (E if cfg.hn_single_head else D, H)
    s["output_projection.1.bias"] = (E if cfg.hn_single_head else D,)
    if cfg.separate_out_embeddings and not cfg.hn_single_head:
        projector("output_projection_out.0.")
        s["output_projection_out.1.weight"] = (D, H)
        s["output_projection_out.1.bias"] = (D,)
    if cfg.hn_rescale_embeddings:
        s["in_scaler.w"] = (1, E)
        s["in_scaler.b"] = (1, E)
        s["scaler.w"] = (1, D)
        s["scaler.b"] = (1, D)
        if cfg.separate_out_embeddings:
            s["out_scaler.w"] = (1, D)
            s["out_scaler.b"] = (1, D)
    if cfg.hn_predict_bias:
        s["bias_projection.weight"] = (1, H)
        s["bias_projection.bias"] = (1,)
    if cfg.hn_embed_lang_id:
        s["lang_embeddings.weight"] = (cfg.n_langs, H)
    return s
