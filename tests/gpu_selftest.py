"""GPU self-test driver (run in a child process so that a kernel fault cannot poison the caller's CUDA context).

    python tests/gpu_selftest.py gemm    --impl {2,3,5}
    python tests/gpu_selftest.py forward --impl {2,3,5} [--configs tiny,tiny_lang,...] [--terms {1,2,3}]

Prints one JSON object per line: GEMM cases are checked against a float64 torch matmul of the SAME fp32 inputs,
forward cases against the numpy oracle (oracle/hypernet_oracle.py).  ``tests/test_gpu_*.py`` assert on the lines.
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import hypernet_oracle as ho  # noqa: E402
import zett_synthetic as synthetic
from zett_b200 import _lib  # noqa: E402
from zett_b200.modeling_hypernet import NativeHypernet  # noqa: E402

GEMM_CASES = [
    # m, n, k, act, terms
    (128, 128, 64, 0, 3),
    (128, 256, 128, 0, 3),
    (256, 256, 256, 0, 3),
    (200, 384, 192, 0, 3),      # ragged M, N = 3 x 128
    (77, 64, 128, 1, 3),        # tiny M, gelu tanh
    (130, 72, 144, 0, 3),       # K and N not multiples of 32: padded lines, partial accumulator chunk
    (1000, 768, 768, 2, 3),     # XLM-R shapes, gelu erf
    (4096, 2304, 768, 0, 3),
    (3000, 4096, 4096, 0, 3),   # Mistral H x H
    (1024, 8192, 4096, 1, 3),
    (1024, 4096, 8192, 0, 3),
    (512, 256, 128, 0, 1),      # single-pass mode
    (2048, 4096, 4096, 0, 1),
    (128, 128, 64, 0, 2),       # fp16 + two e5m2 correction passes
    (200, 384, 192, 0, 2),
    (130, 72, 144, 2, 2),
    (1000, 768, 768, 2, 2),
    (3000, 4096, 4096, 0, 2),
    (1024, 8192, 4096, 1, 2),
    (1024, 4096, 8192, 0, 2),
    (16384, 4096, 4096, 0, 3),  # throughput probes at the last-layer shapes
    (16384, 4096, 4096, 0, 2),
]

# whole-epilogue cases (zett_gemm_f32_ex): m, n, k, act, terms, residual, affine, operand output
# M >= 16384, N % 512 == 0, K >= 4096 fill whole waves of 256 x 512 tiles, so gemm_impl 5 takes its wide branch
GEMM_EX_CASES = [
    (300, 256, 128, 1, 2, True, True, True),
    (300, 256, 128, 2, 3, True, True, True),
    (130, 72, 144, 1, 2, True, False, True),
    (16384, 4096, 4096, 1, 2, True, False, True),
    (16384, 4096, 4096, 2, 2, False, True, True),
    (16384, 8192, 4096, 2, 2, False, False, True),
    (16384, 4096, 8192, 0, 2, True, True, False),
    (16384, 4096, 4096, 1, 3, True, True, True),
    (18944, 1024, 4096, 2, 2, True, True, True),
]


def _gemm_ex(lib, a, w, b, res, cs, ct, out, out_op, act, impl, terms, iters=0, report=False):
    m, k = a.shape
    n = w.shape[0]
    ms = ctypes.c_float(0)
    buf = ctypes.create_string_buffer(1024) if report else None
    ptr = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())  # noqa: E731
    _lib.check(lib.zett_gemm_f32_ex(ptr(a), ptr(w), ptr(b), ptr(res), ptr(cs), ptr(ct), ptr(out), ptr(out_op), m, n, k, act, impl,
                                    terms, max(iters, 1), ctypes.byref(ms) if iters else None, buf, 1024 if report else 0, None))
    rep = json.loads(buf.value.decode()) if (report and buf.value) else None
    return ms.value, rep


def _ref_gemm(a, w, b, act, res=None, cs=None, ct=None):
    ref = a.double() @ w.double().T
    if b is not None:
        ref = ref + b.double()
    if act == 1:
        ref = torch.nn.functional.gelu(ref, approximate="tanh")
    elif act == 2:
        ref = torch.nn.functional.gelu(ref)
    if res is not None:
        ref = ref + res.double()
    if cs is not None:
        ref = cs.double() * ref + ct.double()
    return ref


def run_gemm(impl: int):
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    ok = True
    tol = {3: 5e-5, 2: 2e-4, 1: 2e-2}
    for (m, n, k, act, terms) in GEMM_CASES:
        a = torch.randn(m, k, device=dev)
        w = torch.randn(n, k, device=dev) / k ** 0.5
        b = torch.randn(n, device=dev) * 0.1
        out = torch.full((m, n), float("nan"), device=dev)
        iters = 3
        try:
            ms, _ = _gemm_ex(lib, a, w, b, None, None, None, out, None, act, impl, terms, iters=iters)
        except Exception as e:  # noqa: BLE001
            print(json.dumps(dict(kind="gemm", impl=impl, m=m, n=n, k=k, act=act, terms=terms, error=str(e))), flush=True)
            return False
        ref = _ref_gemm(a, w, b, act)
        err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
        fro = ((out.double() - ref).norm() / ref.norm()).item()
        good = bool(np.isfinite(err) and err < tol[terms])
        ok &= good
        tflops = 2.0 * m * n * k * iters / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        print(json.dumps(dict(kind="gemm", impl=impl, m=m, n=n, k=k, act=act, terms=terms, max_rel=err, fro_rel=fro,
                              ok=good, ms_per_launch=ms / iters, tflops=tflops)), flush=True)
        del a, w, out, ref
    for (m, n, k, act, terms, use_res, use_aff, use_op) in GEMM_EX_CASES:
        a = torch.randn(m, k, device=dev)
        w = torch.randn(n, k, device=dev) / k ** 0.5
        b = torch.randn(n, device=dev) * 0.1
        res = torch.randn(m, n, device=dev) if use_res else None
        cs = (0.5 + torch.rand(n, device=dev)) if use_aff else None
        ct = (torch.randn(n, device=dev) * 0.02) if use_aff else None
        out = torch.full((m, n), float("nan"), device=dev)
        out_op = torch.full((m, n), float("nan"), device=dev) if use_op else None
        try:
            _gemm_ex(lib, a, w, b, res, cs, ct, out, out_op, act, impl, terms)
        except Exception as e:  # noqa: BLE001
            print(json.dumps(dict(kind="gemm_ex", impl=impl, m=m, n=n, k=k, act=act, terms=terms, error=str(e))), flush=True)
            return False
        ref = _ref_gemm(a, w, b, act, res, cs, ct)
        scale = ref.abs().max().item()
        err = (out.double() - ref).abs().max().item() / scale
        line = dict(kind="gemm_ex", impl=impl, m=m, n=n, k=k, act=act, terms=terms, residual=use_res, affine=use_aff,
                    operand_out=use_op, max_rel=err)
        good = bool(np.isfinite(err) and err < tol[terms])
        if use_op:
            # the operand lines written for the next GEMM decode to the fp32 output within the format's own resolution
            # (fp16 + e5m2 residual: ~2^-13 of the value; bf16 hi + lo: ~2^-16)
            d = (out_op.double() - out.double()).abs()
            op_rel = (d / (out.double().abs() + 1e-3 * scale)).max().item()
            line["operand_vs_f32"] = op_rel
            good &= bool(np.isfinite(op_rel) and op_rel < (4e-4 if terms == 2 else 5e-5))
        line["ok"] = good
        ok &= good
        print(json.dumps(line), flush=True)
        del a, w, out, ref, res, out_op
    return ok


def run_sweep(shapes=None, impls=(2, 5), terms_list=(2, 3, 1), acts=(0, 0x100, 2, "res", "erf_op", "tanh_op")):
    """Timing probes of the GEMM engine (no parity claim): ms per launch over shapes x formats x tile shapes, with the
    per-role stall picture of one launch when ZETT_GEMM_PROF=1."""
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    shapes = shapes or [(16384, 4096, 4096), (53248, 12288, 4096), (53248, 4096, 8192), (53248, 8192, 4096), (54000, 2304, 768),
                        (54000, 1536, 768), (54000, 768, 1536), (53248, 6144, 2048), (53248, 4096, 2048)]
    for (m, n, k) in shapes:
        a = torch.randn(m, k, device=dev)
        w = torch.randn(n, k, device=dev) / k ** 0.5
        b = torch.randn(n, device=dev) * 0.1
        out = torch.empty((m, n), device=dev)
        for impl in impls:
            for terms in terms_list:
                for act in acts:
                    iters = 4
                    res, o32, oop, label = None, out, None, act
                    if act == "res":   # fp32 output + residual (attention-output / MLP-down epilogue)
                        res, act = torch.randn(m, n, device=dev), 0
                    elif act in ("erf_op", "tanh_op"):   # GELU + operand-line output only (MLP-up / ProjectorBlock dense1 epilogue)
                        o32, oop, act = None, out, (2 if act == "erf_op" else 1)
                    ms, rep = _gemm_ex(lib, a, w, b, res, None, None, o32, oop, act, impl, terms, iters=iters, report=True)
                    act = label
                    del res
                    t = ms / iters
                    print(json.dumps(dict(kind="sweep", m=m, n=n, k=k, impl=impl, terms=terms, act=act, ms=round(t, 4),
                                          tflops=round(2.0 * m * n * k / (t * 1e-3) / 1e12, 1), prof=rep)), flush=True)
        del a, w, out
    return True


class PowerSampler:
    """nvidia-smi power / SM clock every 100 ms (GPU 0)."""

    def __init__(self):
        import subprocess
        import tempfile
        fd, self.path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        self.proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw,temperature.gpu",
                                      "--format=csv,noheader,nounits", "-lms", "100"], stdout=open(self.path, "w"),
                                     stderr=subprocess.DEVNULL)

    def stop(self):
        self.proc.terminate()
        self.proc.wait(timeout=5)
        rows = []
        for ln in open(self.path):
            try:
                rows.append([float(x) for x in ln.split(",")])
            except ValueError:
                pass
        os.unlink(self.path)
        rows = rows[len(rows) // 3:]  # steady state: drop the ramp
        if not rows:
            return {}
        a = np.array(rows)
        return dict(sm_mhz=float(np.median(a[:, 0])), power_w=float(np.median(a[:, 1])), temp_c=float(a[:, 2].max()), samples=len(a))


def run_sustained(mnk_list, seconds=1.5):
    """Energy / throughput probes (no parity claim): every operand format and cluster shape launched back to back for
    ~`seconds`, with board power and SM clock sampled meanwhile (act 0x100 = no stores, to separate the cost of the epilogue's
    output from the main loop).  A cuBLAS bf16 matmul of the same shape runs beside them."""
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    for (m, n, k) in mnk_list:
        a = torch.randn(m, k, device=dev)
        w = torch.randn(n, k, device=dev) / k ** 0.5
        out = torch.empty((m, n), device=dev)
        # (label, impl, terms, store the output?, environment)
        W_LAST, STREAM = {"ZETT_L2_HINT_W": "3"}, {"ZETT_STREAM_OUT": "1"}
        variants = [
            ("f16+2xe5m2 256x256", 2, 2, True, {}),
            ("f16+2xe5m2 256x256 no store", 2, 2, False, {}),
            ("f16+2xe5m2 256x512", 5, 2, True, {}),
            ("f16+2xe5m2 256x512 no store", 5, 2, False, {}),
            ("f16+2xe5m2 256x512 W=evict_last, streaming stores", 5, 2, True, {**W_LAST, **STREAM}),
            ("bf16x3 256x256", 2, 3, True, {}),
            ("bf16x3 256x512", 5, 3, True, {}),
            ("bf16 single pass 256x256", 2, 1, True, {}),
            ("bf16 single pass 256x256 no store", 2, 1, False, {}),
            ("bf16 single pass 256x512", 5, 1, True, {}),
        ]
        only = os.environ.get("ZETT_SUSTAINED_ONLY")  # comma-separated substrings of the labels to keep
        if only:
            variants = [v for v in variants if any(tag in v[0] for tag in only.split(","))]
        b = torch.zeros(n, device=dev)
        for (label, impl, terms, store, env) in variants:
            for kk in ("ZETT_L2_HINT_W", "ZETT_L2_HINT_A", "ZETT_STREAM_OUT"):
                os.environ.pop(kk, None)
            os.environ.update(env)
            act = 0 if store else 0x100
            ms, _ = _gemm_ex(lib, a, w, b, None, None, None, out, None, act, impl, terms, iters=2)
            iters = max(4, int(seconds * 1e3 / max(ms / 2, 1e-3)))
            ps = PowerSampler()
            ms, _ = _gemm_ex(lib, a, w, b, None, None, None, out, None, act, impl, terms, iters=iters)
            st = ps.stop()
            t = ms / iters
            print(json.dumps(dict(kind="sustained", label=label, m=m, n=n, k=k, impl=impl, terms=terms,
                                  store=store, iters=iters, ms=round(t, 4),
                                  tflops_once=round(2.0 * m * n * k / (t * 1e-3) / 1e12, 1), **st)), flush=True)
        for kk in ("ZETT_L2_HINT_W", "ZETT_L2_HINT_A", "ZETT_STREAM_OUT"):
            os.environ.pop(kk, None)
        ab, wb = a.bfloat16(), w.bfloat16()
        ob = torch.empty((m, n), device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            torch.matmul(ab, wb.T, out=ob)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            torch.matmul(ab, wb.T, out=ob)
        e1.record()
        torch.cuda.synchronize()
        iters = max(4, int(seconds * 1e3 / (e0.elapsed_time(e1) / 4)))
        ps = PowerSampler()
        e0.record()
        for _ in range(iters):
            torch.matmul(ab, wb.T, out=ob)
        e1.record()
        torch.cuda.synchronize()
        st = ps.stop()
        t = e0.elapsed_time(e1) / iters
        print(json.dumps(dict(kind="sustained", m=m, n=n, k=k, impl="cublas_bf16", iters=iters, ms=round(t, 4),
                              tflops_once=round(2.0 * m * n * k / (t * 1e-3) / 1e12, 1), **st)), flush=True)
        del a, w, out, ab, wb, ob
    return True


RASTER_KEYS = ("ZETT_RASTER_CHUNK_MB", "ZETT_RASTER_GROUP_M", "ZETT_L2_HINT_A", "ZETT_L2_HINT_W")


def run_raster(m, n, k, combos, impl=5, terms=2, seconds=1.2):
    """Rasterisation / L2-hint probes of one GEMM shape under sustained load (no parity claim): `combos` is a list of
    (chunk MB, m-group, hint A, hint W); hints 1 evict_first, 2 normal, 3 evict_last.  The same combinations go through
    `one` under ncu for their DRAM traffic (scripts/gpu_r2k.sh)."""
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    a = torch.randn(m, k, device=dev)
    w = torch.randn(n, k, device=dev) / k ** 0.5
    out = torch.empty((m, n), device=dev)
    b = torch.zeros(n, device=dev)
    for combo in combos:
        for key, val in zip(RASTER_KEYS, combo):
            os.environ[key] = str(val)
        ms, _ = _gemm_ex(lib, a, w, b, None, None, None, out, None, 0, impl, terms, iters=2)
        iters = max(4, int(seconds * 1e3 / max(ms / 2, 1e-3)))
        ps = PowerSampler()
        ms, _ = _gemm_ex(lib, a, w, b, None, None, None, out, None, 0, impl, terms, iters=iters)
        st = ps.stop()
        t = ms / iters
        print(json.dumps(dict(kind="raster", combo=list(combo), m=m, n=n, k=k, impl=impl, terms=terms, iters=iters, ms=round(t, 4),
                              tflops_once=round(2.0 * m * n * k / (t * 1e-3) / 1e12, 1), **st)), flush=True)
    for key in RASTER_KEYS:
        os.environ.pop(key, None)
    return True


def run_one(m, n, k, impl, terms):
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    a = torch.randn(m, k, device=dev)
    w = torch.randn(n, k, device=dev) / k ** 0.5
    b = torch.randn(n, device=dev) * 0.1
    out = torch.empty((m, n), device=dev)
    ms, rep = _gemm_ex(lib, a, w, b, None, None, None, out, None, 0, impl, terms, iters=1, report=True)
    print(json.dumps(dict(kind="one", m=m, n=n, k=k, impl=impl, terms=terms, ms=ms, prof=rep)), flush=True)
    return True


def forward_case(name, impl, rows, lang, overrides=None, max_rows_per_pass=0, seed=13, terms=0):
    dev = torch.device("cuda", 0)
    cfg = synthetic.make_config(name, **(overrides or {}))
    weights = synthetic.make_weights(cfg, seed=11)
    big = cfg.original_vocab_size > 100000
    src = synthetic.make_source_embeddings(cfg, seed=12)
    sf = synthetic.make_random_surface_forms(cfg, rows, seed=seed)
    t0 = time.time()
    want = ho.hypernet_forward(cfg, weights, sf, src, lang_index=lang)
    t_oracle = time.time() - t0
    nat = NativeHypernet(cfg, weights, dev, max_rows_per_pass=max_rows_per_pass, gemm_impl=impl, split_terms=terms)
    sf_d = torch.from_numpy(sf).to(dev)
    src_d = torch.from_numpy(src).to(dev)
    D = cfg.n_embd
    pred_in = torch.full((rows, D), float("nan"), device=dev)
    pred_out = torch.full((rows, D), float("nan"), device=dev) if cfg.separate_out_embeddings else None
    bias = torch.full((rows,), float("nan"), device=dev)
    res = dict(kind="forward", config=name, impl=impl, rows=rows, overrides=overrides or {}, pass_rows=max_rows_per_pass,
               oracle_s=round(t_oracle, 3), big_table=big)
    try:
        nat.forward_into(sf_d, src_d, -1 if lang is None else lang, pred_in, pred_out, bias)
        nat.check()
        torch.cuda.synchronize()
        t0 = time.time()
        nat.forward_into(sf_d, src_d, -1 if lang is None else lang, pred_in, pred_out, bias)
        nat.check()
        res["gpu_s"] = round(time.time() - t0, 4)
    except Exception as e:  # noqa: BLE001
        res["error"] = str(e)
        print(json.dumps(res), flush=True)
        return False
    masked = ho.fully_masked_rows(cfg, sf)
    ok = True
    for key, g, w in (("pred_in", pred_in, want[0]), ("pred_out", pred_out, want[1]), ("pred_bias", bias, want[2])):
        if w is None:
            continue
        gn = g.cpu().numpy()
        fro, worst = ho.rel_errors(gn, w, exclude=masked)
        res[key] = [fro, worst]
        ok &= bool(np.isfinite(gn).all() and fro < 1e-3 and worst < 1e-3)
        if masked.any():
            mfro, mworst = ho.rel_errors(gn[masked], w[masked])
            res[key + "_masked_rows"] = [mfro, mworst]
    res["stats"] = nat.stats()
    res["ok"] = ok
    print(json.dumps(res), flush=True)
    nat.close()
    return ok


FORWARD_CASES = {
    # name: (config, rows, lang, overrides)
    "tiny": ("tiny", 64, None, None),
    "tiny_lang": ("tiny_lang", 64, 3, None),
    "tiny_single_head": ("tiny", 48, None, {"hn_single_head": True}),
    "tiny_plain": ("tiny", 48, None, {"hn_rescale_embeddings": False, "hn_predict_bias": False,
                                       "separate_out_embeddings": False, "hn_n_extra_tokens": 0}),
    "tiny_one_layer": ("tiny", 40, None, {"hn_n_layers": 1}),
    "tiny_multi_pass": ("tiny", 300, None, None),
    "xlmr": ("xlmr", 512, 3, None),
    "tinyllama": ("tinyllama", 256, None, None),
    "mistral": ("mistral", 192, None, None),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["gemm", "forward", "sweep", "one", "sustained", "raster"])
    ap.add_argument("--combos", default="48,4,2,2", help="raster: semicolon-separated chunkMB,groupM,hintA,hintW")
    ap.add_argument("--mnk", default="16384,4096,4096")
    ap.add_argument("--impl", type=int, default=0)
    ap.add_argument("--configs", default="tiny,tiny_lang,tiny_single_head,tiny_plain,tiny_one_layer,tiny_multi_pass")
    ap.add_argument("--terms", type=int, default=0)
    ap.add_argument("--sweep-terms", default="2,3,1")
    args = ap.parse_args()
    if args.what == "one":
        m, n, k = [int(x) for x in args.mnk.split(",")]
        ok = run_one(m, n, k, args.impl or 5, args.terms or 2)
    elif args.what == "sweep":
        shapes = None if args.mnk == "16384,4096,4096" else [tuple(int(x) for x in t.split(",")) for t in args.mnk.split(";")]
        ok = run_sweep(shapes, terms_list=tuple(int(x) for x in args.sweep_terms.split(",")))
    elif args.what == "raster":
        m, n, k = [int(x) for x in args.mnk.split(",")]
        ok = run_raster(m, n, k, [tuple(int(x) for x in c.split(",")) for c in args.combos.split(";")], args.impl or 5, args.terms or 2)
    elif args.what == "sustained":
        ok = run_sustained([tuple(int(x) for x in t.split(",")) for t in args.mnk.split(";")])
    elif args.what == "gemm":
        ok = run_gemm(args.impl or 5)
    else:
        ok = True
        for c in args.configs.split(","):
            name, rows, lang, ov = FORWARD_CASES[c]
            ok &= forward_case(name, args.impl, rows, lang, ov, max_rows_per_pass=96 if c == "tiny_multi_pass" else 0,
                               terms=args.terms)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
