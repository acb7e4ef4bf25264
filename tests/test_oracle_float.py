"""The numpy oracle (oracle/hypernet_oracle.py) against goldens minted from the reference's own
PyTorch hypernet (tests/golden/make_golden.py)."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import hypernet_oracle as ho
import zett_synthetic as synthetic

CASES = sorted(os.path.basename(p)[len("hypernet_"):-4]
               for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "hypernet_*.npz")))


def load_case(golden_dir, case):
    g = np.load(os.path.join(golden_dir, f"hypernet_{case}.npz"))
    meta = json.loads(str(g["meta"]))
    cfg = synthetic.make_config(meta["config"], **meta["overrides"])
    weights = synthetic.make_weights(cfg, seed=meta["weight_seed"])
    src = synthetic.make_source_embeddings(cfg, seed=meta["source_seed"])
    checksum = float(sum(v.astype(np.float64).sum() for v in weights.values()) + src.astype(np.float64).sum())
    assert abs(checksum - meta["input_checksum"]) < 1e-6 * max(1.0, abs(checksum)), "synthetic generator drifted"
    return g, meta, cfg, weights, src


def test_golden_cases_present():
    assert {"tiny", "tiny_lang", "tiny_single_head", "tiny_plain"} <= set(CASES)


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference(golden_dir, case):
    g, meta, cfg, weights, src = load_case(golden_dir, case)
    sf = g["surface_forms"]
    pred_in, pred_out, bias = ho.hypernet_forward(cfg, weights, sf, src, lang_index=meta["lang_index"])
    masked = ho.fully_masked_rows(cfg, sf)
    # fp32 numpy vs fp32 torch differ only by summation order: 1e-5 relative is ~100x tighter than the
    # 1e-3 budget of the CUDA path.  Fully masked rows follow the eager uniform-softmax semantics too.
    for name, got in (("pred_in", pred_in), ("pred_out", pred_out), ("pred_bias", bias)):
        if name not in g.files:
            assert got is None or name == "pred_bias"
            continue
        fro, worst = ho.rel_errors(got, g[name])
        assert fro < 1e-5 and worst < 2e-5, (case, name, fro, worst)
    assert masked.sum() == (0 if cfg.hn_embed_lang_id else 1)


def test_oracle_fp64_agrees_with_fp32(golden_dir):
    g, meta, cfg, weights, src = load_case(golden_dir, "tiny")
    a = ho.hypernet_forward(cfg, weights, g["surface_forms"], src, dtype=np.float64)
    fro, worst = ho.rel_errors(a[0], g["pred_in"])
    assert fro < 1e-5 and worst < 2e-5


def test_row_independence(golden_dir):
    """Each row's result is independent of the rest of the batch: permutation invariance."""
    g, meta, cfg, weights, src = load_case(golden_dir, "tiny")
    sf = g["surface_forms"]
    perm = np.random.default_rng(0).permutation(len(sf))
    a = ho.hypernet_forward(cfg, weights, sf, src)
    b = ho.hypernet_forward(cfg, weights, sf[perm][:17], src)
    np.testing.assert_allclose(b[0], a[0][perm][:17], rtol=0, atol=2e-6)


def test_unsupported_branches_raise():
    cfg = synthetic.make_config("tiny", hn_add_inter_token_attention=True)
    with pytest.raises(NotImplementedError):
        ho.hypernet_forward(cfg, {}, np.zeros((1, 7), np.int32), np.zeros((300, 128), np.float32))
    cfg = synthetic.make_config("tiny", hn_model_type="t5")
    with pytest.raises(NotImplementedError):
        ho.hypernet_forward(cfg, {}, np.zeros((1, 7), np.int32), np.zeros((300, 128), np.float32))


@pytest.mark.parametrize("case", CASES)
def test_torch_oracle_matches_reference(golden_dir, case):
    """The threaded torch-CPU restatement (bench.py's CPU baseline) against the same reference-minted goldens."""
    import torch
    from oracle import hypernet_oracle_torch as hot
    g, meta, cfg, weights, src = load_case(golden_dir, case)
    out = hot.hypernet_forward(cfg, hot.to_torch(weights), g["surface_forms"], torch.from_numpy(src),
                               lang_index=meta["lang_index"])
    for name, got in zip(("pred_in", "pred_out", "pred_bias"), out):
        if name not in g.files:
            continue
        fro, worst = ho.rel_errors(got, g[name])
        assert fro < 1e-5 and worst < 2e-5, (case, name, fro, worst)
