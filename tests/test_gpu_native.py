"""GPU parity tests proper (run with -m gpu on a B200): every compute call goes through the C ABI of
libzett_b200.so and is compared with the oracle (oracle/) or a float64 torch matmul.

The per-implementation sweeps run in child processes (tests/gpu_selftest.py) so that a fault in one kernel variant
is reported as such instead of poisoning the CUDA context of the whole session."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def run_selftest(args, timeout=900):
    r = subprocess.run([sys.executable, os.path.join(HERE, "gpu_selftest.py")] + args, capture_output=True, text=True,
                       timeout=timeout, cwd=ROOT)
    lines = []
    for ln in r.stdout.splitlines():
        ln = ln.strip()
        if ln.startswith("{"):
            lines.append(json.loads(ln))
    return r, lines


@pytest.mark.parametrize("impl", [3, 2, 5])
def test_gemm_engine(impl):
    """tcgen05 CTA pairs on 256 x 256 tiles / on 256 x 512 tiles, and the SIMT checker, against a float64 matmul: every
    operand format, ragged and padded sizes, and the whole epilogue (GELU, residual, column affine, operand-line output)
    on shapes that take the wide-tile branch (M >= 16384, N % 512 == 0, K >= 4096)."""
    r, lines = run_selftest(["gemm", "--impl", str(impl)])
    assert lines, r.stdout + r.stderr
    for ln in lines:
        assert "error" not in ln, ln
        assert ln["ok"], ln
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("impl", [3, 2, 5])
def test_forward_tiny_configs(impl):
    """All tiny configurations (separate / tied heads, lang-id slot, single head, no rescale / bias, one layer,
    multi-pass) against the oracle: Frobenius and worst-row relative error <= 1e-3."""
    r, lines = run_selftest(["forward", "--impl", str(impl)])
    assert len(lines) == 6, r.stdout + r.stderr
    for ln in lines:
        assert "error" not in ln, ln
        assert ln["ok"], ln


@pytest.mark.parametrize("config", ["xlmr", "tinyllama", "mistral"])
def test_forward_baseline_shapes(config):
    """The three BASELINE shapes on a few hundred rows against the oracle (default GEMM implementation)."""
    r, lines = run_selftest(["forward", "--impl", "0", "--configs", config], timeout=1800)
    assert len(lines) == 1, r.stdout + r.stderr
    assert "error" not in lines[0], lines[0]
    assert lines[0]["ok"], lines[0]


# ---- in-process tests of the public Python surface (default implementation) ---------------------------------------
@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    return torch


def _model(torch, name, seed=11, **overrides):
    import zett_synthetic as synthetic
    from zett_b200.modeling_hypernet import ZettHypernet, load_weights_numpy
    cfg = synthetic.make_config(name, **overrides)
    weights = synthetic.make_weights(cfg, seed=seed)
    model = load_weights_numpy(ZettHypernet(cfg), weights).to("cuda")
    return cfg, weights, model


def test_golden_fixtures_through_public_api(torch_cuda, golden_dir):
    """tests/golden/hypernet_*.npz were minted by the reference itself; the CUDA path must match them to 1e-3."""
    import glob
    torch = torch_cuda
    from oracle import hypernet_oracle as ho
    import zett_synthetic as synthetic
    for path in sorted(glob.glob(os.path.join(golden_dir, "hypernet_*.npz"))):
        g = np.load(path)
        meta = json.loads(str(g["meta"]))
        cfg, weights, model = _model(torch, meta["config"], seed=meta["weight_seed"], **meta["overrides"])
        src = synthetic.make_source_embeddings(cfg, seed=meta["source_seed"])
        sf = g["surface_forms"]
        lang = meta["lang_index"]
        out = model(torch.from_numpy(sf).cuda(), source_embeddings=torch.from_numpy(src).cuda(),
                    lang_index=None if lang is None else torch.tensor(lang))
        masked = ho.fully_masked_rows(cfg, sf)
        for name, got in zip(("pred_in", "pred_out", "pred_bias"), out):
            if name not in g.files:
                assert got is None or name == "pred_bias"
                continue
            fro, worst = ho.rel_errors(got.cpu().numpy(), g[name], exclude=masked)
            assert fro < 1e-3 and worst < 1e-3, (path, name, fro, worst)


def test_row_independence_and_idempotence(torch_cuda):
    """Permutation / batch-composition invariance and run-to-run determinism, bit-exact."""
    torch = torch_cuda
    import zett_synthetic as synthetic
    cfg, weights, model = _model(torch, "tiny")
    src = torch.from_numpy(synthetic.make_source_embeddings(cfg, seed=12)).cuda()
    sf = synthetic.make_random_surface_forms(cfg, 500, seed=3)
    a = model(torch.from_numpy(sf).cuda(), source_embeddings=src)
    b = model(torch.from_numpy(sf).cuda(), source_embeddings=src)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    perm = np.random.default_rng(0).permutation(len(sf))
    c = model(torch.from_numpy(sf[perm][:123]).cuda(), source_embeddings=src)
    for x, y in zip(a, c):
        assert torch.equal(x[torch.from_numpy(perm[:123]).cuda()], y)
    model.max_rows_per_pass = 64
    model.refresh()
    d = model(torch.from_numpy(sf).cuda(), source_embeddings=src)
    for x, y in zip(a, d):
        assert torch.equal(x, y)


def test_out_of_range_id_raises(torch_cuda):
    torch = torch_cuda
    import zett_synthetic as synthetic
    cfg, weights, model = _model(torch, "tiny")
    src = torch.from_numpy(synthetic.make_source_embeddings(cfg, seed=12)).cuda()
    sf = synthetic.make_random_surface_forms(cfg, 16, seed=3)
    sf[7, 1] = cfg.original_vocab_size + cfg.hn_n_extra_tokens  # first id past the fallback table
    with pytest.raises(IndexError):
        model(torch.from_numpy(sf).cuda(), source_embeddings=src)
    sf[7, 1] = -1
    with pytest.raises(IndexError):
        model(torch.from_numpy(sf).cuda(), source_embeddings=src)


def test_unsupported_branches_raise(torch_cuda):
    torch = torch_cuda
    import zett_synthetic as synthetic
    from zett_b200.modeling_hypernet import ZettHypernet
    with pytest.raises(NotImplementedError):
        ZettHypernet(synthetic.make_config("tiny", hn_add_inter_token_attention=True))
    with pytest.raises(NotImplementedError):
        ZettHypernet(synthetic.make_config("tiny", hn_model_type="t5"))
    cfg, weights, model = _model(torch, "tiny")
    sf = torch.zeros((2, 7), dtype=torch.int32).cuda()
    with pytest.raises(NotImplementedError):
        model(sf, target_priors=torch.zeros(2), source_embeddings=torch.zeros(300, 128).cuda())


def test_end_to_end_tokens_to_embeddings(torch_cuda):
    """token strings -> native retokenizer -> H2D -> forward -> D2H through transfer.make_predict / batched_inference,
    against the oracle on the same surface forms; batched_inference == one-shot prediction."""
    torch = torch_cuda
    from oracle import hypernet_oracle as ho
    import zett_synthetic as synthetic
    from zett_b200.surface_forms import get_surface_form_matrix
    from zett_b200.transfer import batched_inference, default_args, make_predict
    hn = synthetic.make_hn_tokenizer("unigram", 316, seed=5, pad_token="</s>")  # ids < tiny's V0 + n_extra = 316
    tokens = synthetic.make_target_tokens(700, seed=6)
    cfg, weights, model = _model(torch, "tiny", pad_token_id=int(hn.pad_token_id))
    sf, n_trunc = get_surface_form_matrix(tokens, cfg.hn_surface_maxlen, hn)
    assert sf.max() < cfg.original_vocab_size + cfg.hn_n_extra_tokens
    src = synthetic.make_source_embeddings(cfg, seed=12)
    predict = make_predict(model, src)
    one = predict(sf)
    want = ho.hypernet_forward(cfg, weights, sf, src)
    masked = ho.fully_masked_rows(cfg, sf)
    for g, w in zip(one, want):
        fro, worst = ho.rel_errors(g, w, exclude=masked)
        assert fro < 1e-3 and worst < 1e-3
    cfg.hidden_size = cfg.n_embd
    bi = batched_inference(sf, None, cfg, default_args(batch_size=256), predict, rng=np.random.default_rng(1))
    for g, w in zip(bi, one):
        np.testing.assert_array_equal(g, w)


def test_full_vocab_properties_xlmr(torch_cuda):
    """BASELINE config 2 at full size (50 257 rows, XLM-R shape): size-independent properties -- finite outputs,
    duplicate surface-form rows give bit-identical predictions, pass-size invariance on a slice."""
    torch = torch_cuda
    import zett_synthetic as synthetic
    cfg, weights, model = _model(torch, "xlmr")
    src = torch.from_numpy(synthetic.make_source_embeddings(cfg, seed=12)).cuda()
    sf = synthetic.make_random_surface_forms(cfg, 50257, seed=5)
    sf[40000:40100] = sf[100:200]
    out = model(torch.from_numpy(sf).cuda(), source_embeddings=src, lang_index=torch.tensor(3))
    assert out[1] is None
    assert torch.isfinite(out[0]).all() and torch.isfinite(out[2]).all()
    assert torch.equal(out[0][40000:40100], out[0][100:200])
    assert torch.equal(out[2][40000:40100], out[2][100:200])
    st = model.native().stats()
    assert st["rows"] == 50257 and st["kernel_launches"] > 0 and st["packed_positions"] > 0


def test_pipelined_tokens_path_equals_one_shot(torch_cuda):
    """transfer.predict_from_tokens (retokenise / H2D / compute / D2H overlapped per pass) == one-shot prediction,
    bit for bit, including the surface forms it builds on the way."""
    torch = torch_cuda
    import zett_synthetic as synthetic
    from zett_b200.surface_forms import get_surface_form_matrix
    from zett_b200.transfer import TokenPipeline, make_predict, predict_from_tokens
    hn = synthetic.make_hn_tokenizer("unigram", 316, seed=5, pad_token="</s>")
    tokens = synthetic.make_target_tokens(1500, seed=6)
    cfg, weights, model = _model(torch, "tiny", pad_token_id=int(hn.pad_token_id))
    src = torch.from_numpy(synthetic.make_source_embeddings(cfg, seed=12)).cuda()
    sf, n_trunc = get_surface_form_matrix(tokens, cfg.hn_surface_maxlen, hn)
    one = make_predict(model, src)(sf)
    piped = predict_from_tokens(model, tokens, hn, src, rows_per_pass=256)
    for a, b in zip(one, piped):
        np.testing.assert_array_equal(a, b)
    pipe = TokenPipeline(model, hn, src, rows_per_pass=400)
    out, sf2, nt2 = pipe.run(tokens)
    np.testing.assert_array_equal(sf2, sf)
    assert nt2 == n_trunc
    np.testing.assert_array_equal(out[:, :cfg.n_embd].numpy(), one[0])
    # the first pass cut in two (the GPU starts after `first_chunk_rows` tokens): same rows, same bits
    pipe = TokenPipeline(model, hn, src, rows_per_pass=400, first_chunk_rows=96)
    out, sf2, nt2 = pipe.run(tokens)
    np.testing.assert_array_equal(sf2, sf)
    assert nt2 == n_trunc
    np.testing.assert_array_equal(out[:, :cfg.n_embd].numpy(), one[0])


@pytest.mark.parametrize("terms", [2, 3])
def test_explicit_operand_formats_parity(terms):
    """Both multi-term operand formats, requested explicitly, meet the 1e-3 budget: split_terms = 2 (fp16 MMA + two
    e5m2 correction MMAs through kind::f8f6f4; the default when the sizes allow it) and 3 (three bf16 MMA terms)."""
    r, lines = run_selftest(["forward", "--impl", "2", "--terms", str(terms)])
    assert len(lines) == 6, r.stdout + r.stderr
    for ln in lines:
        assert "error" not in ln and ln["ok"], ln


def test_shape_variants(torch_cuda):
    """Paths the shipped configs do not reach: 2 layers, maxlen 12 (rows longer than 8 positions), head sizes 32 and
    128, a vocabulary of one row and an empty one."""
    torch = torch_cuda
    from oracle import hypernet_oracle as ho
    import zett_synthetic as synthetic
    # n_embd = 72: GEMM K = 144 and N = 72 are not multiples of 32 (padded operand lines, partial accumulator chunks)
    for overrides in (dict(hn_n_layers=2, hn_surface_maxlen=12), dict(hn_num_attention_heads=4), dict(hn_num_attention_heads=1),
                      dict(n_embd=72)):
        cfg, weights, model = _model(torch, "tiny", **overrides)
        src_np = synthetic.make_source_embeddings(cfg, seed=12)
        sf = synthetic.make_random_surface_forms(cfg, 77, seed=9)
        got = model(torch.from_numpy(sf).cuda(), source_embeddings=torch.from_numpy(src_np).cuda())
        want = ho.hypernet_forward(cfg, weights, sf, src_np)
        masked = ho.fully_masked_rows(cfg, sf)
        for g, w in zip(got, want):
            fro, worst = ho.rel_errors(g.cpu().numpy(), w, exclude=masked)
            assert fro < 1e-3 and worst < 1e-3, (overrides, fro, worst)
    cfg, weights, model = _model(torch, "tiny")
    src = torch.from_numpy(synthetic.make_source_embeddings(cfg, seed=12)).cuda()
    sf = synthetic.make_random_surface_forms(cfg, 9, seed=9)
    full = model(torch.from_numpy(sf).cuda(), source_embeddings=src)
    one = model(torch.from_numpy(sf[4:5]).cuda(), source_embeddings=src)
    for a, b in zip(full, one):
        assert torch.equal(a[4:5], b)
    empty = model(torch.zeros((0, 7), dtype=torch.int32).cuda(), source_embeddings=src)
    assert empty[0].shape == (0, cfg.n_embd) and empty[2].shape == (0,)


def test_full_vocab_properties_mistral(torch_cuda):
    """BASELINE config 4 at full size (50 304 rows, Mistral-7B shape): finite, duplicate rows identical, a slice
    recomputed alone reproduces the whole-vocabulary result bit for bit (pass-composition invariance), and 96 rows taken
    out of the full-size run meet the 1e-3 budget against the fp32 oracle."""
    torch = torch_cuda
    import zett_synthetic as synthetic
    from oracle import hypernet_oracle as ho
    from zett_b200.modeling_hypernet import NativeHypernet
    cfg = synthetic.make_config("mistral")
    weights = synthetic.make_weights(cfg, seed=0)
    nat = NativeHypernet(cfg, weights, torch.device("cuda", 0))
    src_np = synthetic.make_source_embeddings(cfg, seed=100)
    src = torch.from_numpy(src_np).cuda()
    sf = synthetic.make_random_surface_forms(cfg, 50304, seed=5)
    sf[30000:30064] = sf[64:128]
    sf_d = torch.from_numpy(sf).cuda()
    D = cfg.n_embd
    outs = [torch.empty((50304, D), device="cuda"), torch.empty((50304, D), device="cuda"), torch.empty((50304,), device="cuda")]
    nat.forward_into(sf_d, src, -1, *outs)
    nat.check()
    for o in outs:
        assert torch.isfinite(o).all()
        assert torch.equal(o[30000:30064], o[64:128])
    part = [torch.empty((300, D), device="cuda"), torch.empty((300, D), device="cuda"), torch.empty((300,), device="cuda")]
    nat.forward_into(sf_d[20000:20300].contiguous(), src, -1, *part)
    nat.check()
    for o, p in zip(outs, part):
        assert torch.equal(o[20000:20300], p)
    st = nat.stats()
    assert st["rows"] == 300 and st["distinct_ids"] > 0
    pick = np.arange(40000, 40096)
    want = ho.hypernet_forward(cfg, weights, sf[pick], src_np)
    masked = ho.fully_masked_rows(cfg, sf[pick])
    for o, w in zip(outs, want):
        fro, worst = ho.rel_errors(o[40000:40096].cpu().numpy(), w, exclude=masked)
        assert fro < 1e-3 and worst < 1e-3, (fro, worst)
    nat.close()


@pytest.mark.parametrize("name,lang", [("tiny", None), ("tiny_lang", 2)])
def test_pair_dedup_is_bit_exact(torch_cuda, name, lang, monkeypatch):
    """The first encoder layer evaluated once per distinct (id, position) pair must reproduce the per-position
    evaluation bit for bit (same arithmetic on the same operands), with and without the lang-id slot, over several
    passes; the statistics report fewer pairs than positions on a vocabulary with repeated pieces."""
    torch = torch_cuda
    import zett_synthetic as synthetic
    from zett_b200.modeling_hypernet import NativeHypernet
    cfg = synthetic.make_config(name)
    weights = synthetic.make_weights(cfg, seed=11)
    src = torch.from_numpy(synthetic.make_source_embeddings(cfg, seed=12)).cuda()
    sf = synthetic.make_random_surface_forms(cfg, 700, seed=21)
    sf[:, 0] = sf[:, 0] % 7  # few distinct (id, position 0) pairs
    sf_d = torch.from_numpy(sf).cuda()
    D = cfg.n_embd
    results = []
    for flag in ("1", "0"):
        monkeypatch.setenv("ZETT_DEDUP_PAIRS", flag)
        nat = NativeHypernet(cfg, weights, torch.device("cuda", 0), max_rows_per_pass=256)
        outs = [torch.full((700, D), float("nan"), device="cuda"),
                torch.full((700, D), float("nan"), device="cuda") if cfg.separate_out_embeddings else None,
                torch.full((700,), float("nan"), device="cuda")]
        nat.forward_into(sf_d, src, -1 if lang is None else lang, *outs)
        nat.check()
        st = nat.stats()
        if flag == "1":
            assert 0 < st["distinct_pairs"] < st["encoder_positions"], st
        else:
            assert st["distinct_pairs"] == 0, st
        results.append(outs)
        nat.close()
    for a, b in zip(*results):
        if a is not None:
            assert torch.isfinite(a).all()
            assert torch.equal(a, b)


@pytest.mark.parametrize("case", ["weights_x1e3", "weights_x1e-4", "embeddings_x1e3", "heavy_tailed_ln", "mixed_row_scales"])
def test_operand_range_guard(torch_cuda, case):
    """The fp16 + e5m2 operand format has five exponent bits; the fp32 reference (hf_hypernet/modeling_hypernet.py:179-189)
    has eight.  Weights are stored with a power-of-two scale per output row, and a forward that produces an ACTIVATION
    outside fp16's range is detected on the device (zett_hn_stats.operand_overflows, ZETT_ERR_RANGE) and repeated with the
    three-term bf16 split -- so every case below meets the 1e-3 budget without the caller doing anything."""
    import warnings
    torch = torch_cuda
    from oracle import hypernet_oracle as ho
    import zett_synthetic as synthetic
    from zett_b200.modeling_hypernet import ZettHypernet, load_weights_numpy
    cfg = synthetic.make_config("tiny")
    weights = synthetic.make_weights(cfg, seed=11)
    src_np = synthetic.make_source_embeddings(cfg, seed=12)
    rng = np.random.default_rng(5)
    expect_fallback = False
    if case == "weights_x1e3":
        weights = {k: (v * np.float32(1e3) if k.endswith(".weight") and v.ndim == 2 and "embeddings" not in k and "LayerNorm" not in k else v)
                   for k, v in weights.items()}
        expect_fallback = True   # ProjectorBlock intermediates reach ~1e6
    elif case == "weights_x1e-4":
        weights = {k: (v * np.float32(1e-4) if k.endswith(".weight") and v.ndim == 2 and "embeddings" not in k and "LayerNorm" not in k else v)
                   for k, v in weights.items()}
    elif case == "embeddings_x1e3":
        src_np = src_np * np.float32(1e3)
    elif case == "heavy_tailed_ln":
        for k in list(weights):
            if ("LayerNorm.weight" in k or k.endswith("ln.weight")):
                g = weights[k].copy()
                idx = rng.choice(g.size, size=max(1, g.size // 32), replace=False)
                g[idx] *= rng.uniform(20, 60, size=idx.size).astype(np.float32)
                weights[k] = g
    elif case == "mixed_row_scales":
        for k in list(weights):
            v = weights[k]
            # three decades between the output rows of every Linear except the attention projections: scaling query / key
            # rows makes the softmax itself ill-conditioned (the fp32 and fp64 oracles then differ by 1e-4, a hundred times
            # their usual distance), which says nothing about operand formats
            if k.endswith(".weight") and v.ndim == 2 and "embeddings" not in k and "LayerNorm" not in k and "attention.self" not in k:
                s = (10.0 ** rng.uniform(-2, 1, size=(v.shape[0], 1))).astype(np.float32)
                weights[k] = v * s
    sf = synthetic.make_random_surface_forms(cfg, 96, seed=9)
    model = load_weights_numpy(ZettHypernet(cfg), weights).to("cuda")
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        got = model(torch.from_numpy(sf).cuda(), source_embeddings=torch.from_numpy(src_np).cuda())
    want = ho.hypernet_forward(cfg, weights, sf, src_np)
    masked = ho.fully_masked_rows(cfg, sf)
    for g, w in zip(got, want):
        fro, worst = ho.rel_errors(g.cpu().numpy(), w, exclude=masked)
        assert fro < 1e-3 and worst < 1e-3, (case, fro, worst)
    st = model.native().stats()
    fell_back = any("split_terms = 3" in str(c.message) for c in caught)
    assert st["split_terms"] == (3 if fell_back else 2), (case, st)
    if expect_fallback:
        assert fell_back, case
    # an explicit split_terms = 2 does not fall back: it raises, like an IndexError would
    if fell_back:
        from zett_b200 import _lib
        strict = load_weights_numpy(ZettHypernet(cfg), weights)
        strict.split_terms = 2
        strict = strict.to("cuda")
        with pytest.raises(_lib.OperandRangeError):
            strict(torch.from_numpy(sf).cuda(), source_embeddings=torch.from_numpy(src_np).cuda())


def test_fully_masked_rows_match_the_eager_reference(torch_cuda):
    """A row whose ids are all pad has every key masked; the reference's eager attention (additive finfo.min mask) then
    attends uniformly over all S positions.  The kernels reproduce that: such rows meet the same 1e-3 budget
    as every other row (the bias head, a single dot product that may nearly cancel, is held against the scale of the
    whole bias vector)."""
    torch = torch_cuda
    from oracle import hypernet_oracle as ho
    import zett_synthetic as synthetic
    cfg, weights, model = _model(torch, "tiny")
    src_np = synthetic.make_source_embeddings(cfg, seed=12)
    sf = synthetic.make_random_surface_forms(cfg, 64, seed=17)
    sf[::4, :] = cfg.pad_token_id      # every fourth row fully masked
    got = model(torch.from_numpy(sf).cuda(), source_embeddings=torch.from_numpy(src_np).cuda())
    want = ho.hypernet_forward(cfg, weights, sf, src_np)
    masked = ho.fully_masked_rows(cfg, sf)
    assert masked.sum() >= 16
    for name, g, w in zip(("pred_in", "pred_out", "pred_bias"), got, want):
        g = g.cpu().numpy()
        if name == "pred_bias":
            err = np.abs(g[masked] - w[masked]).max() / np.sqrt(np.mean(w.astype(np.float64) ** 2))
            assert err < 1e-3, (name, err)
        else:
            fro, worst = ho.rel_errors(g[masked], w[masked])
            assert fro < 1e-3 and worst < 1e-3, (name, fro, worst)
