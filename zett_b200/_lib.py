"""ctypes binding of ``libzett_b200.so`` (C ABI in ``include/zett_b200.h``) and its in-tree build.

There is no CPU fallback: if the shared library is missing ``load()`` raises, and every compute entry point needs a
Blackwell GPU (the library refuses other devices).
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
import threading
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "lib", "libzett_b200.so")
SOURCES = ["hypernet.cu", "retok.cpp", "comm.cpp", "sampler.cpp"]
HEADERS = ["ptx.cuh", "operand.cuh", "gemm_tcgen05.cuh", "epilogue.cuh", "kernels.cuh", "unicode_tables.inc"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--shared",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-ldl"]

ZETT_OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_STATE, ERR_INDEX, ERR_KEY, ERR_MISSING_UNK, ERR_RANGE = 0, -1, -2, -3, -4, -5, -6, -7, -8
ABI_VERSION = 2
F32, F16, BF16 = 0, 1, 2


class ZettHnConfig(ctypes.Structure):
    """``zett_hn_config`` (include/zett_b200.h)."""
    _fields_ = [(n, c_int32) for n in (
        "struct_bytes", "hn_surface_maxlen", "hn_n_layers", "n_embd", "hn_hidden_size", "hn_intermediate_size",
        "hn_num_attention_heads", "hn_rescale_embeddings", "hn_embed_target_priors", "hn_add_inter_token_attention",
        "hn_embed_using_source_embeddings", "hn_concat_last_hidden_state", "hn_single_head", "hn_predict_bias",
        "hn_embed_lang_id", "hn_model_type_is_roberta", "n_langs", "pad_token_id", "original_vocab_size",
        "hn_n_extra_tokens", "separate_out_embeddings", "max_position_embeddings")] + [
        ("encoder_layer_norm_eps", c_float), ("max_rows_per_pass", c_int32), ("gemm_impl", c_int32),
        ("split_terms", c_int32)]


class ZettHnStats(ctypes.Structure):
    """``zett_hn_stats`` (include/zett_b200.h)."""
    _fields_ = [("kernel_launches", c_int64), ("rows", c_int64), ("packed_positions", c_int64),
                ("encoder_positions", c_int64), ("flops_executed", c_double), ("gemm_ms", c_double),
                ("gemm_launches", c_int64), ("distinct_ids", c_int64), ("distinct_pairs", c_int64),
                ("split_terms", c_int64), ("gemm_impl", c_int64), ("operand_overflows", c_int64)]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(ROOT, "include", "zett_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA kernels + C ABI for sm_100a into ``zett_b200/lib/libzett_b200.so`` (nvcc cross-compiles
    without a GPU).  The .so stays in-tree so it travels with the repository snapshot."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libzett_b200.so")
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    tmp = LIB_PATH + ".tmp.%d" % os.getpid()
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + SOURCES
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    os.replace(tmp, LIB_PATH)
    if verbose:
        print(r.stderr)
    return LIB_PATH


_lock = threading.Lock()
_lib = None

# name -> (restype, argtypes); must list every symbol include/zett_b200.h declares (tests/test_abi.py checks)
SIGNATURES = {
    "zett_last_error": (c_char_p, []),
    "zett_abi_version": (c_int, []),
    "zett_hn_create": (c_int, [POINTER(ZettHnConfig), POINTER(c_void_p)]),
    "zett_hn_set_weight": (c_int, [c_void_p, c_char_p, c_void_p, c_int, c_int, POINTER(c_int64)]),
    "zett_hn_finalize": (c_int, [c_void_p]),
    "zett_hn_workspace_bytes": (c_size_t, [c_void_p, c_int64]),
    "zett_hn_forward": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p,
                                c_int64, c_int64, c_void_p]),
    "zett_hn_check": (c_int, [c_void_p, c_void_p]),
    "zett_hn_set_split_terms": (c_int, [c_void_p, c_int]),
    "zett_hn_get_stats": (c_int, [c_void_p, POINTER(ZettHnStats)]),
    "zett_hn_set_timing": (c_int, [c_void_p, c_int]),
    "zett_hn_destroy": (None, [c_void_p]),
    "zett_gemm_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_int,
                              c_int, POINTER(c_float), c_void_p]),
    "zett_gemm_f32_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                 c_int64, c_int, c_int, c_int, c_int, POINTER(c_float), c_char_p, c_int64, c_void_p]),
    "zett_comm_unique_id": (c_int, [c_void_p]),
    "zett_comm_init": (c_int, [c_int, c_int, c_void_p, POINTER(c_void_p)]),
    "zett_allgather_rows": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "zett_comm_ipc_handle": (c_int, [c_void_p, c_void_p, POINTER(c_int64)]),
    "zett_comm_register": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, POINTER(c_int64)]),
    "zett_comm_unregister": (c_int, [c_void_p]),
    "zett_comm_barrier": (c_int, [c_void_p, c_void_p]),
    "zett_comm_info": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "zett_comm_destroy": (None, [c_void_p]),
    "zett_tok_create_unigram": (c_int, [POINTER(c_char_p), POINTER(c_double), c_int64, c_int64, c_int, POINTER(c_void_p)]),
    "zett_tok_create_bpe": (c_int, [POINTER(c_char_p), c_int64, POINTER(c_int32), c_int64, c_int64, c_char_p, c_char_p,
                                    c_int, c_int, c_int, POINTER(c_void_p)]),
    "zett_tok_tokenize": (c_int64, [c_void_p, c_char_p, POINTER(c_int32), c_int64]),
    "zett_surface_forms": (c_int, [c_void_p, POINTER(c_char_p), c_int64, POINTER(c_int32), c_int32, c_int32, c_int64,
                                   POINTER(c_int32), POINTER(c_int64), c_int]),
    "zett_surface_forms_blob": (c_int, [c_void_p, c_char_p, c_int64, c_int64, c_char_p, c_int64, POINTER(c_int32), c_int64,
                                        c_int32, c_int32, c_int64, POINTER(c_int32), POINTER(c_int64), c_int]),
    "zett_tok_destroy": (None, [c_void_p]),
    "zett_sampler_create": (c_int, [POINTER(c_void_p)]),
    "zett_sampler_sample": (c_int, [c_void_p, c_char_p, c_int64, POINTER(ctypes.c_uint32), c_int64, c_int64, c_int64, c_int64, c_double,
                                    ctypes.c_uint64, c_int, c_int, POINTER(c_void_p), POINTER(c_int64), POINTER(c_void_p), POINTER(c_int64)]),
    "zett_sampler_free": (None, [c_void_p]),
    "zett_sampler_destroy": (None, [c_void_p]),
}


def load():
    """Load the shared library (building it first when it is missing and nvcc is available)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        stale = os.path.exists(LIB_PATH) and needs_build() and bool(shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"))
        if not os.path.exists(LIB_PATH) or stale:  # sources newer than the .so: rebuild where a compiler exists
            try:
                build(force=stale)
            except Exception as e:  # noqa: BLE001
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        "libzett_b200.so is not built and could not be built here (%s). zett_b200 has no CPU or "
                        "PyTorch fallback: run `python -c 'import __graft_entry__ as g; g.build()'` where nvcc "
                        "is available." % e) from e
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.zett_abi_version() != ABI_VERSION:
            raise RuntimeError("libzett_b200.so ABI version mismatch: rebuild it (python -c 'import __graft_entry__ as g; g.build()')")
        _lib = lib
        return lib


class OperandRangeError(ArithmeticError):
    """A GEMM operand left fp16's range under split_terms = 2 (``ZETT_ERR_RANGE``): the forward has to be repeated with
    the three-term bf16 split.  The wrappers in ``modeling_hypernet`` / ``transfer`` / ``parallel`` do that themselves."""


_EXC = {ERR_RANGE: OperandRangeError, ERR_INVALID: ValueError, ERR_UNSUPPORTED: NotImplementedError, ERR_CUDA: RuntimeError, ERR_STATE: RuntimeError,
        ERR_INDEX: IndexError, ERR_KEY: KeyError, ERR_MISSING_UNK: Exception}


def check(rc: int):
    """Map a negative ``zett_status`` to the Python exception type the reference raises in the same situation."""
    if rc >= 0:
        return rc
    msg = (load().zett_last_error() or b"").decode("utf-8", "replace")
    raise _EXC.get(rc, RuntimeError)(msg or "zett_b200 error %d" % rc)
