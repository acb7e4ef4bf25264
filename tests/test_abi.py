"""The C-ABI shared library loads without a GPU and exports every symbol include/zett_b200.h declares; the ctypes
mirror of the config struct has the C layout; error paths that need no device behave as documented."""
import ctypes
import os
import re
import subprocess
import tempfile

import pytest

from zett_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "zett_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zett_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib.SIGNATURES, "ctypes signature missing for " + n
    assert lib.zett_abi_version() == _lib.ABI_VERSION


def test_struct_layout_matches_c():
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "zett_b200.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu\n", sizeof(zett_hn_config), offsetof(zett_hn_config, encoder_layer_norm_eps),
             offsetof(zett_hn_config, split_terms), sizeof(zett_hn_stats), offsetof(zett_hn_stats, gemm_ms));
      return 0;
    }'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        vals = [int(x) for x in subprocess.check_output([exe]).split()]
    C, S = _lib.ZettHnConfig, _lib.ZettHnStats
    assert vals == [ctypes.sizeof(C), C.encoder_layer_norm_eps.offset, C.split_terms.offset, ctypes.sizeof(S), S.gemm_ms.offset]


def test_unsupported_config_is_rejected_before_any_device_work():
    import zett_synthetic as synthetic
    from zett_b200.modeling_hypernet import make_c_config
    lib = _lib.load()
    h = ctypes.c_void_p()
    for bad in (dict(hn_add_inter_token_attention=True), dict(hn_embed_target_priors=True), dict(hn_model_type="t5"),
                dict(hn_concat_last_hidden_state=True), dict(hn_embed_using_source_embeddings=False)):
        cfg = synthetic.make_config("tiny", **bad)
        rc = lib.zett_hn_create(ctypes.byref(make_c_config(cfg)), ctypes.byref(h))
        assert rc == _lib.ERR_UNSUPPORTED, bad
        with pytest.raises(NotImplementedError):
            _lib.check(rc)
    c = make_c_config(synthetic.make_config("tiny"))
    c.struct_bytes = 8
    assert lib.zett_hn_create(ctypes.byref(c), ctypes.byref(h)) == _lib.ERR_INVALID


def test_no_cpu_fallback():
    """Without a GPU the product path fails loudly instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import zett_synthetic as synthetic
    from zett_b200.modeling_hypernet import ZettHypernet, make_c_config
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.zett_hn_create(ctypes.byref(make_c_config(synthetic.make_config("tiny"))), ctypes.byref(h))
    assert rc == _lib.ERR_CUDA
    model = ZettHypernet(synthetic.make_config("tiny"))
    with pytest.raises(RuntimeError):
        model(torch.zeros((2, 7), dtype=torch.int32), source_embeddings=torch.zeros((300, 128)))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "zett_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
