"""zett_b200 -- B200-native embedding-prediction hot path of ZeTT (Zero-Shot Tokenizer Transfer).

Public surface (mirrors the reference):
  ZettHypernetConfig, ZettHypernet            hf_hypernet/{configuration,modeling}_hypernet.py
  get_surface_form_matrix                     zett/utils.py:651-689
  make_predict, batched_inference             scripts/transfer.py:54-124,221-234
  predict_sharded                             row sharding over GPUs + one all-gather (zett/utils.py:26)
"""
from .config import ZettHypernetConfig  # noqa: F401

__all__ = ["ZettHypernetConfig", "ZettHypernet", "get_surface_form_matrix", "make_predict", "batched_inference",
           "predict_sharded"]


def __getattr__(name):  # torch-dependent parts load lazily so that the config / generators import fast
    if name == "ZettHypernet":
        from .modeling_hypernet import ZettHypernet
        return ZettHypernet
    if name == "register_auto_classes":
        from .modeling_hypernet import register_auto_classes
        return register_auto_classes
    if name == "get_surface_form_matrix":
        from .surface_forms import get_surface_form_matrix
        return get_surface_form_matrix
    if name in ("make_predict", "batched_inference", "predict_whole_vocab"):
        from . import transfer
        return getattr(transfer, name)
    if name == "predict_sharded":
        from .parallel import predict_sharded
        return predict_sharded
    raise AttributeError(name)
