"""``ZettHypernetConfig`` -- same field names and defaults as the reference's config
(hf_hypernet/configuration_hypernet.py:3-56) so a trained checkpoint's ``config.json`` loads unchanged.

Fields that training writes onto the config without declaring them (SURVEY.md section 8b, config contract):
``original_vocab_size`` (train.py:314), ``hn_n_extra_tokens`` (train.py:361), ``separate_out_embeddings``
(train.py:350), ``pad_token_id`` (train.py:295), ``langs`` (scripts/convert_to_pt.py:31-33) ride in through
``**kwargs`` exactly as they do in the reference.
"""
from __future__ import annotations

from transformers import PretrainedConfig

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
