#!/usr/bin/env python
"""Benchmark of the ZeTT embedding-prediction hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config mistral|tinyllama|xlmr]

A "step" predicts the embeddings of one whole synthetic vocabulary (rows_per_gpu rows on every GPU, weak scaling):
``value`` times the forward with the surface-form matrix already resident in HBM (CUDA events, max over ranks);
``e2e`` times the user-facing path from token strings on the host: native retokenizer -> pinned H2D -> forward ->
single all-gather -> pinned D2H of the predicted matrices.  ``--impl reference`` times the reference's own CPU
execution of the path (threaded ATen restatement in oracle/) on a bounded sample of the same workload.
Prints ONE JSON line on rank 0.
"""
import argparse
import csv
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "predicted token-embeddings/sec (whole vocab)"
WORKLOADS = {
    # BASELINE.json configs[3] / [2] / [1]; rows = GPT-NeoX vocab padded to 128 (50 277 -> 50 304) or GPT-2 (50 257)
    "mistral": dict(workload="Mistral-7B hypernet (d=4096), GPT-NeoX 50k vocab", rows=50304, hn="bpe", lang=None),
    "tinyllama": dict(workload="TinyLlama-1.1B hypernet (d=2048), GPT-NeoX 50k vocab", rows=50304, hn="bpe", lang=None),
    "xlmr": dict(workload="xlm-roberta-base hypernet (d=768), GPT2 50k vocab", rows=50257, hn="unigram", lang=3),
}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(tflops=float(p.get("bf16_tflops_sustained", p.get("bf16_tflops"))), hbm=float(p["hbm_gbs"]), src="measured (sustained)")
    return dict(tflops=1400.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        with open(self.path) as f:
            for line in f:
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 8 or parts[0] != str(self.gpu):
                    continue
                try:
                    sm.append(float(parts[1])); mx.append(float(parts[2])); power.append(float(parts[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        os.unlink(self.path)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def build_workload(name, rank, rows):
    import zett_synthetic as synthetic
    wl = WORKLOADS[name]
    cfg = synthetic.make_config(name)
    hn = synthetic.make_hn_tokenizer(wl["hn"], 32000, seed=1, pad_token="</s>" if cfg.pad_token_id == 2 else "<pad>",
                                     fit_total=True)
    assert hn.pad_token_id == cfg.pad_token_id, (hn.pad_token_id, cfg.pad_token_id)
    tokens = synthetic.make_target_tokens(rows, seed=2 + rank)
    return cfg, hn, tokens


# ------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's CPU execution of the path on host cores, bounded sample
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_run(name, steps, warmup, budget_s):
    import torch
    from oracle import hypernet_oracle_torch as hot
    from oracle import retok_oracle as ro
    import zett_synthetic as synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = WORKLOADS[name]
    cfg, hn, tokens = build_workload(name, 0, wl["rows"])
    W = hot.to_torch(synthetic.make_weights(cfg, seed=0))
    src = torch.from_numpy(synthetic.make_source_embeddings(cfg, seed=100))

    def one(sample_tokens):
        sf, _ = ro.surface_form_matrix_hf(sample_tokens, cfg.hn_surface_maxlen, hn)
        out = hot.hypernet_forward(cfg, W, sf, src, lang_index=wl["lang"])
        return out

    # calibrate the sample so that (steps + warmup) samples fit the budget
    n0 = 64
    t0 = time.perf_counter(); one(tokens[300:300 + n0]); dt = time.perf_counter() - t0
    t0 = time.perf_counter(); one(tokens[300:300 + n0]); dt = min(dt, time.perf_counter() - t0)
    per_row = dt / n0
    total = steps + warmup
    n = int(max(64, min(wl["rows"], budget_s / max(total, 1) / per_row)))
    n = max(64, (n // 64) * 64)
    sample = tokens[300:300 + n] if 300 + n <= len(tokens) else tokens[:n]
    for _ in range(warmup):
        one(sample)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter(); one(sample); times.append(time.perf_counter() - t0)
    mean = float(np.mean(times))
    return dict(value=n / mean, ms_per_step=mean * 1e3, cores=cores, rows=n,
                sample="%d of %d rows per step (rows are independent; rows/s extrapolates linearly); HF tokenizers loop + "
                       "threaded ATen fp32 forward" % (n, wl["rows"]))


def torch_eager_gpu_run(name, dev, rows=8192):
    """The library-kernel bar: the reference's own operator sequence through PyTorch eager (ATen / cuBLAS
    sm_100 kernels) on the same B200, fp32 and with TF32 matmuls enabled.  Rows are independent, so a sample is timed."""
    import torch
    from oracle import hypernet_oracle_torch as hot
    import zett_synthetic as synthetic
    wl = WORKLOADS[name]
    cfg = synthetic.make_config(name)
    W = hot.to_torch(synthetic.make_weights(cfg, seed=0), dev)
    src = torch.from_numpy(synthetic.make_source_embeddings(cfg, seed=100)).to(dev)
    sf = synthetic.make_random_surface_forms(cfg, rows, seed=7)
    out = {}
    for label, tf32 in (("fp32", False), ("tf32", True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        for _ in range(2):
            hot.hypernet_forward(cfg, W, sf, src, lang_index=wl["lang"])
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            hot.hypernet_forward(cfg, W, sf, src, lang_index=wl["lang"])
        e1.record()
        torch.cuda.synchronize(dev)
        out[label] = rows * 3 / (e0.elapsed_time(e1) * 1e-3)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    del W, src
    torch.cuda.empty_cache()
    return {"unit": "rows/s", "rows_sampled": rows, "fp32": out["fp32"], "tf32": out["tf32"],
            "what": "reference operator sequence via torch eager (ATen/cuBLAS) on the same GPU, dense over all L positions"}


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    r = cpu_reference_run(args.config, args.steps, args.warmup, budget_s=150.0)
    wl = WORKLOADS[args.config]
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "rows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["workload"], "rows_per_step": r["rows"], "device": "host CPU"},
        "cpu_baseline": {"value": r["value"], "unit": "rows/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def profile_of_largest_gemm(config, fmt_tag, impl):
    """DRAM bytes (read + write) of one launch of the largest GEMM of the step, from the committed ncu summary."""
    if config != "mistral":
        return None
    tags = ([fmt_tag + "_wide"] if impl == 5 else []) + [fmt_tag]  # 256 x 512 pair tiles have their own capture
    for path in [os.path.join(ROOT, "profiles", "gemm_tcgen05_%s_%s_53248x12288x4096.csv" % (rnd, tag)) for rnd in ("r1",) for tag in tags]:
        if not os.path.exists(path):
            continue
        vals = {}
        with open(path) as f:
            for row in csv.reader(f):
                if len(row) >= 3:
                    vals[row[0]] = (row[1], row[2])
        try:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "byte": 1.0, "Kbyte": 1e3}
            rd = float(vals["dram__bytes_read.sum"][1]) * scale[vals["dram__bytes_read.sum"][0]]
            wr = float(vals["dram__bytes_write.sum"][1]) * scale[vals["dram__bytes_write.sum"][0]]
            act = vals.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", ("%", "nan"))[1]
        except (KeyError, ValueError):
            continue
        return {"traffic": int(rd + wr),
                "note": "%s: one isolated launch of the QKV GEMM of a 53248-position pass under ncu --set full; "
                        "sm__pipe_tensor_cycles_active %s %%; traffic = its dram read + write bytes" % (os.path.relpath(path, ROOT), act)}
    return None


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import zett_synthetic as synthetic
from zett_b200 import parallel
    from zett_b200.modeling_hypernet import NativeHypernet
    from zett_b200.surface_forms import get_surface_form_matrix

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200 GPU: zett_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = WORKLOADS[args.config]
    rows = args.rows or wl["rows"]
    cfg, hn, tokens = build_workload(args.config, rank, rows)
    lang = wl["lang"]
    weights = synthetic.make_weights(cfg, seed=0)
    nat = NativeHypernet(cfg, weights, dev, max_rows_per_pass=args.rows_per_pass, gemm_impl=args.gemm_impl,
                         split_terms=args.split_terms)  # C-ABI handle
    del weights
    src = torch.from_numpy(synthetic.make_source_embeddings(cfg, seed=100)).to(dev)
    D, separate = cfg.n_embd, bool(cfg.separate_out_embeddings)
    width = parallel.packed_width(D, separate)
    lang_i = -1 if lang is None else lang

    sf_host, n_trunc = get_surface_form_matrix(tokens, cfg.hn_surface_maxlen, hn)
    hist = synthetic.length_histogram(sf_host, cfg.pad_token_id)
    sf_dev = torch.from_numpy(sf_host).to(dev)
    block = torch.zeros((rows, width), dtype=torch.float32, device=dev)
    full = torch.empty((world * rows, width), dtype=torch.float32, device=dev) if world > 1 else block

    def forward_resident():
        nat.forward_into(sf_dev, src, lang_i, block[:, 0:], block[:, D:] if separate else None,
                         block[:, (2 if separate else 1) * D:], ld_pred=width, ld_bias=width)
        if world > 1:
            dist.all_gather_into_tensor(full, block)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ("value") ----------------------------------------------------------------------
    nat.set_timing(True)
    for _ in range(max(args.warmup, 3)):
        forward_resident()
    nat.check()
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gemm_ms, gemm_flops, launches = 0.0, 0.0, 0
    sync_all()
    e0.record()
    for _ in range(args.steps):
        forward_resident()
    e1.record()
    sync_all()
    elapsed_ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    nat.check()
    st = nat.stats()  # statistics of the last step
    gemm_ms, gemm_flops, launches = st["gemm_ms"], st["flops_executed"], st["kernel_launches"]
    ms_per_step = elapsed_ms / args.steps
    value = world * rows / (ms_per_step * 1e-3)
    terms = int(st["split_terms"])
    fmt = {3: ("bf16x3->f32", "bf16x3", 3.0), 2: ("f16+2xe5m2->f32", "f16f8", 2.0), 1: ("bf16->f32", "bf16", 1.0)}[terms]

    # ---- end to end from host token strings ("e2e") -----------------------------------------------------------------
    nat.set_timing(False)
    from zett_b200.transfer import TokenPipeline
    pipe = TokenPipeline(nat, hn, src, lang, rows_per_pass=args.rows_per_pass or 16384)

    def gather(blk):
        if world > 1:
            dist.all_gather_into_tensor(full, blk)                               # the single collective

    def e2e_step():
        # host token strings -> native retokenizer -> pinned H2D -> forward -> (all-gather) -> pinned D2H of this rank's
        # rows; passes are pipelined (host retokenisation and D2H of pass k overlap the compute of pass k + 1)
        return pipe.run(tokens, after_compute=gather)

    e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    sync_all()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / args.steps
    e2e_value = world * rows / e2e_s
    out_pinned, sf_e2e, _ = e2e_step()
    assert np.isfinite(out_pinned.numpy()[:, : (2 if separate else 1) * D + 1]).all()
    assert np.array_equal(sf_e2e, sf_host)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    f_ref = cfg.flops_per_row(pruned=False)
    line = {
        "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": fmt[0], "data": "synthetic",
        "config": {
            "workload": wl["workload"], "rows_per_gpu": rows, "total_rows": world * rows, "parallelism": "rows x%d" % world,
            "hn_tokenizer": wl["hn"] + " 32k synthetic", "nonpad_length_histogram": hist, "truncated": n_trunc,
            "l2": "inputs larger than L2 (weights + per-pass activations are GBs; nothing is re-read from a warm L2 by design)",
            "gemm_impl": int(st["gemm_impl"]), "split_terms": terms, "rows_per_pass": args.rows_per_pass or 16384,
            "distinct_ids_per_step": int(st["distinct_ids"]), "distinct_id_position_pairs_per_step": int(st["distinct_pairs"]),
            "packed_positions_per_step": int(st["encoder_positions"]),
            "dense_gflop_per_row_reference": f_ref / 1e9,
            "executed_gflop_per_row": gemm_flops / rows / 1e9 if gemm_flops else None,
        },
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "rows/s", "h2d_bytes_per_step": int(world * rows * cfg.hn_surface_maxlen * 4),
                "d2h_bytes_per_step": int(world * rows * width * 4), "ms_per_step": e2e_s * 1e3},
        "gpu_launches": int(launches * args.steps),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                     "frac": (achieved / peaks["tflops"]) if achieved else None, "traffic": None,
                     "kernel": "gemm_tcgen05_kernel", "peak_source": peaks["src"],
                     "note": "achieved = FLOPs of the GEMMs issued in one step (each product counted once although the operand "
                             "format issues several MMA terms per product) / summed CUDA-event time of those launches",
                     "gemm_ms_per_step": gemm_ms, "gemm_launches_per_step": st["gemm_launches"],
                     "mma_issue_tflops_f16_equivalent": (achieved * fmt[2]) if achieved else None},
    }
    prof = profile_of_largest_gemm(args.config, fmt[1], int(st["gemm_impl"]))
    if prof:
        line["roofline"]["traffic"] = prof["traffic"]
        line["roofline"]["profile"] = prof["note"]
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(args.config, 1, 1, budget_s=args.cpu_seconds)
        line["cpu_baseline"] = {"value": r["value"], "unit": "rows/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        try:
            line["torch_eager_b200"] = torch_eager_gpu_run(args.config, dev)
        except Exception as e:  # noqa: BLE001
            line["torch_eager_b200"] = {"error": str(e)[:200]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="mistral", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="rows per GPU (default: the workload's vocabulary size)")
    ap.add_argument("--rows-per-pass", type=int, default=0, help="vocabulary rows per pass of the kernels (0 = the library default)")
    ap.add_argument("--gemm-impl", type=int, default=0)
    ap.add_argument("--split-terms", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=30.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
