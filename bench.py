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
import math
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
    """nvidia-smi clocks / throttle reasons of ONE GPU every 200 ms.  It is started before the warm-up steps (nvidia-smi needs
    about a second to come up on an 8-GPU box, longer than a short timed region) and every sample carries its timestamp, so
    that `stop` can keep exactly the samples taken while the timed region ran."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self, t_begin=None, t_end=None):
        """t_begin / t_end: datetime bounds of the timed region (None = keep every sample)."""
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        rows = []
        with open(self.path) as f:
            for line in f:
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 9 or parts[1] != str(self.gpu):
                    continue
                try:
                    ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f")
                    rows.append((ts, float(parts[2]), float(parts[3]), float(parts[4]), parts[5:9]))
                except ValueError:
                    continue
        os.unlink(self.path)
        inside = [r for r in rows if t_begin is None or (t_begin <= r[0] <= t_end)]
        window = "timed region"
        if not inside and rows:   # region shorter than the sampling period: the samples under load right around it
            inside = [r for r in rows if r[3] > 0.5 * max(x[3] for x in rows)]
            window = "warm-up + timed region (no sample fell inside the timed region itself)"
        reasons = set()
        for r in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm = [r[1] for r in inside]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(r[2] for r in inside) if inside else None,
                "power_w_max": max(r[3] for r in inside) if inside else None, "samples": len(inside), "window": window,
                "reasons": sorted(reasons)}


def build_workload(name, rank, rows, vocab="random", hn_kind=None):
    import zett_synthetic as synthetic
    wl = WORKLOADS[name]
    cfg = synthetic.make_config(name)
    hn = synthetic.make_hn_tokenizer(hn_kind or wl["hn"], 32000, seed=1, pad_token="</s>" if cfg.pad_token_id == 2 else "<pad>",
                                     fit_total=True)
    assert hn.pad_token_id == cfg.pad_token_id, (hn.pad_token_id, cfg.pad_token_id)
    if vocab == "concat":
        tokens = synthetic.make_concat_tokens(rows, hn, seed=3 + rank)
    else:
        tokens = synthetic.make_target_tokens(rows, seed=2 + rank)
    return cfg, hn, tokens


# ------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's CPU execution of the path on host cores, bounded sample
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_run(name, steps, warmup, budget_s):
    """The reference's own modules (hf_hypernet.ZettHypernet, zett.utils.get_surface_form_matrix) staged under
    oracle/_ref by oracle/make_ref.py when they are there (kind "reference"), else the restatement in oracle/ (kind
    "port"): fp32, all host cores, on a bounded sample of the workload."""
    import torch
    from oracle import hypernet_oracle_torch as hot
    from oracle import ref_loader
    from oracle import retok_oracle as ro
    import zett_synthetic as synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = WORKLOADS[name]
    cfg, hn, tokens = build_workload(name, 0, wl["rows"])
    weights = synthetic.make_weights(cfg, seed=0)
    src = torch.from_numpy(synthetic.make_source_embeddings(cfg, seed=100))
    kind = "port"
    if ref_loader.available():
        try:
            model = ref_loader.build_hypernet(cfg, weights)
            ref_utils = ref_loader.load_utils()
            kind = "reference"
        except Exception as e:  # noqa: BLE001
            print("bench: oracle/_ref not usable (%s); timing the port" % str(e)[:200], file=sys.stderr)
    if kind == "reference":
        lang = None if wl["lang"] is None else torch.tensor(wl["lang"])

        def one(sample_tokens):
            sf, _ = ref_utils.get_surface_form_matrix(sample_tokens, cfg.hn_surface_maxlen, hn)
            with torch.no_grad():
                return model(torch.from_numpy(sf), source_embeddings=src, lang_index=lang)
        what = "the reference's own get_surface_form_matrix + hf_hypernet.ZettHypernet (oracle/_ref), fp32, eager attention"
    else:
        W = hot.to_torch(weights)

        def one(sample_tokens):
            sf, _ = ro.surface_form_matrix_hf(sample_tokens, cfg.hn_surface_maxlen, hn)
            return hot.hypernet_forward(cfg, W, sf, src, lang_index=wl["lang"])
        what = "HF tokenizers loop + threaded ATen fp32 forward (restatement in oracle/)"
    del weights

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
    return dict(value=n / mean, ms_per_step=mean * 1e3, cores=cores, rows=n, kind=kind,
                sample="%d of %d rows per step (rows are independent; rows/s extrapolates linearly); %s" % (n, wl["rows"], what))


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
        "cpu_baseline": {"value": r["value"], "unit": "rows/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


PROFILE_ROUND = "r2"


def profile_of_largest_gemm(config, fmt_tag, wide):
    """DRAM bytes (read + write) of one launch of the largest GEMM of the step, from THIS round's committed ncu summary
    (profiles/gemm_tcgen05_r2_*.csv; a capture of another round describes another kernel and is refused)."""
    if config != "mistral":
        return None
    tags = ([fmt_tag + "_wide"] if wide else []) + [fmt_tag]  # 256 x 512 pair tiles have their own capture
    for tag in tags:
        path = os.path.join(ROOT, "profiles", "gemm_tcgen05_%s_%s_53248x12288x4096.csv" % (PROFILE_ROUND, tag))
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
        return {"traffic": int(rd + wr), "profile_round": PROFILE_ROUND,
                "note": "%s: one isolated launch of the QKV GEMM of a 53248-position pass under ncu --set full; "
                        "sm__pipe_tensor_cycles_active %s %%; traffic = its dram read + write bytes" % (os.path.relpath(path, ROOT), act)}
    return None


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
class Dist:
    """Process-group plumbing (torch.distributed over NCCL) + the library's own communicator for the data path."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = env_int("WORLD_SIZE", 1)
        self.rank = env_int("RANK", 0)
        self.local_rank = env_int("LOCAL_RANK", 0)
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a B200 GPU: zett_b200 has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.comm = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            from zett_b200 import parallel
            self.comm = parallel.NativeComm.from_torch_distributed()   # zett_comm_init: ncclCommInitRank under the C ABI
            self.side = torch.cuda.Stream(device=self.dev)

    def sync_all(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.comm is not None:
            self.comm.close()
        if self.world > 1:
            self.dist.destroy_process_group()


def measure(d, name, args, rows=None, vocab="random", hn_kind=None, env=None, want_e2e=True, want_parity=True, steps=None,
            min_region_ms=0.0):
    """One workload on every rank: resident-input throughput (`value`), end-to-end throughput from host token strings
    (`e2e`), the GEMM roofline figure, clocks, and parity of the timed output against the oracle."""
    import torch
    import zett_synthetic as synthetic
    from oracle import hypernet_oracle as ho   # the checker: never on the timed path
    from zett_b200 import parallel
    from zett_b200.modeling_hypernet import NativeHypernet
    from zett_b200.surface_forms import get_surface_form_matrix
    from zett_b200.transfer import TokenPipeline

    world, rank, dev = d.world, d.rank, d.dev
    steps = steps or args.steps
    wl = WORKLOADS[name]
    rows = rows or args.rows or wl["rows"]
    old_env = {}
    for k, v in (env or {}).items():
        old_env[k] = os.environ.get(k)
        os.environ[k] = v
    cfg, hn, tokens = build_workload(name, rank, rows, vocab=vocab, hn_kind=hn_kind)
    lang = wl["lang"]
    weights = synthetic.make_weights(cfg, seed=0)
    nat = NativeHypernet(cfg, weights, dev, max_rows_per_pass=args.rows_per_pass, gemm_impl=args.gemm_impl,
                         split_terms=args.split_terms)  # C-ABI handle
    for k, v in old_env.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    src_np = synthetic.make_source_embeddings(cfg, seed=100)
    src = torch.from_numpy(src_np).to(dev)
    D, separate = cfg.n_embd, bool(cfg.separate_out_embeddings)
    width = parallel.packed_width(D, separate)
    lang_i = -1 if lang is None else lang
    P = args.rows_per_pass or 16384

    sf_host, n_trunc = get_surface_form_matrix(tokens, cfg.hn_surface_maxlen, hn)
    hist = synthetic.length_histogram(sf_host, cfg.pad_token_id)
    sf_dev = torch.from_numpy(sf_host).to(dev)
    # super-blocks of world * P rows: rank r owns the r-th slice of each (zett_b200/parallel.py); every rank holds `rows` rows
    plan = parallel.shard_plan(world * rows, world, P)
    full = torch.zeros((plan[-1][0] + world * plan[-1][1], width), dtype=torch.float32, device=dev)
    main_stream = torch.cuda.current_stream(dev)
    transport = "none"
    if world > 1:
        transport = os.environ.get("ZETT_GATHER", "p2p")
        if transport == "p2p":   # peer copies over NVLink (copy engines) instead of ncclAllGather: the GEMMs keep every SM
            if not d.comm.register(full):
                transport = "nccl"   # a rank could not export / map the buffers: the library's ncclAllGather transport

    def slot_of(base, per, n_here):
        return full[base + rank * per: base + rank * per + n_here]

    def forward_resident():
        loc = 0
        for base, per in plan:
            n_here = min(per, rows - loc)
            slot = slot_of(base, per, n_here)
            nat.forward_into(sf_dev[loc:loc + n_here], src, lang_i, slot[:, 0:], slot[:, D:] if separate else None,
                             slot[:, (2 if separate else 1) * D:], ld_pred=width, ld_bias=width)
            if world > 1:   # in-place all-gather of this super-block on the side stream, under the next one's compute
                ev = torch.cuda.Event()
                ev.record(main_stream)
                d.side.wait_event(ev)
                d.comm.allgather_rows(full[base: base + world * per], per, stream=d.side)
            loc += n_here
        if world > 1:
            main_stream.wait_stream(d.side)
            d.comm.barrier(main_stream)

    # ---- device-resident throughput ("value") ----------------------------------------------------------------------
    import datetime
    nat.set_timing(True)
    sampler = ClockSampler(d.local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(max(args.warmup, 3)):
        if i == max(args.warmup, 3) - 1:
            e0.record()
        forward_resident()
    e1.record()
    nat.check()
    d.sync_all()
    if min_region_ms > 0:
        # the `extra` workloads choose their own step count: long enough for the 200 ms clock sampler to see the region
        steps = max(steps, int(math.ceil(min_region_ms / max(d.max_over_ranks(e0.elapsed_time(e1)), 1e-3))))
    d.sync_all()
    t_begin = datetime.datetime.now()
    e0.record()
    for _ in range(steps):
        forward_resident()
    e1.record()
    d.sync_all()
    t_end = datetime.datetime.now()
    elapsed_ms = d.max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    nat.check()
    st = nat.stats()  # totals over the `steps` timed steps (the library accumulates between two checks)
    ms_per_step = elapsed_ms / steps
    value = world * rows / (ms_per_step * 1e-3)
    terms = int(st["split_terms"])
    fmt = {3: ("bf16x3->f32", "bf16x3", 3.0), 2: ("f16+2xe5m2->f32", "f16f8", 2.0), 1: ("bf16->f32", "bf16", 1.0)}[terms]
    # GEMM time (CUDA events around every GEMM launch INSIDE the timed region), FLOPs and launches of one step
    gemm_ms, gemm_flops = st["gemm_ms"] / steps, st["flops_executed"] / steps
    launches, gemm_launches = st["kernel_launches"] // steps, st["gemm_launches"] // steps
    distinct_ids, distinct_pairs, positions = st["distinct_ids"] // steps, st["distinct_pairs"] // steps, st["encoder_positions"] // steps
    nat.set_timing(False)

    # ---- parity of the TIMED output (<= 1e-3 Frobenius and worst row against the fp32 oracle) ------------
    parity = None
    if want_parity:
        rng = np.random.default_rng(1234 + rank)
        pick = np.sort(rng.choice(rows, size=min(args.parity_rows, rows), replace=False))
        glob = np.empty_like(pick)     # row of `full` holding local row i
        loc = 0
        for base, per in plan:
            n_here = min(per, rows - loc)
            m = (pick >= loc) & (pick < loc + n_here)
            glob[m] = base + rank * per + (pick[m] - loc)
            loc += n_here
        got = full[torch.from_numpy(glob).to(dev)].cpu().numpy()
        want = ho.hypernet_forward(cfg, weights, sf_host[pick], src_np, lang_index=lang)
        masked = ho.fully_masked_rows(cfg, sf_host[pick])
        cols = [("pred_in", got[:, :D], want[0])]
        if separate:
            cols.append(("pred_out", got[:, D:2 * D], want[1]))
        cols.append(("pred_bias", got[:, (2 if separate else 1) * D], want[2]))
        fro = worst = 0.0
        for _, g, w in cols:
            f, wr = ho.rel_errors(g, w, exclude=masked)
            fro, worst = max(fro, f), max(worst, wr)
        parity = {"fro": fro, "worst_row": worst, "rows": int(len(pick)), "budget": 1e-3, "ok": bool(fro < 1e-3 and worst < 1e-3),
                  "against": "oracle/hypernet_oracle.py (numpy fp32 restatement pinned to reference-minted goldens) on rows "
                             "sampled from the timed output"}
        if world > 1:
            # every rank recomputes 64 rows of its NEIGHBOUR's shard and holds them against the gathered matrix: bit-exact
            nb = (rank + 1) % world
            _, _, nb_tokens = build_workload(name, nb, rows, vocab=vocab, hn_kind=hn_kind)
            a = int(rng.integers(0, max(1, rows - 64)))
            nb_sf, _ = get_surface_form_matrix(nb_tokens[a:a + 64], cfg.hn_surface_maxlen, hn)
            blk = torch.zeros((len(nb_sf), width), dtype=torch.float32, device=dev)
            nat.forward_into(torch.from_numpy(nb_sf).to(dev), src, lang_i, blk[:, 0:], blk[:, D:] if separate else None,
                             blk[:, (2 if separate else 1) * D:], ld_pred=width, ld_bias=width)
            nat.check()
            loc, same = 0, None
            for base, per in plan:
                n_here = min(per, rows - loc)
                lo, hi = max(a, loc), min(a + len(nb_sf), loc + n_here)
                if hi > lo:
                    theirs = full[base + nb * per + (lo - loc): base + nb * per + (hi - loc)]
                    ok = bool(torch.equal(theirs[:, :(2 if separate else 1) * D + 1], blk[lo - a:hi - a, :(2 if separate else 1) * D + 1]))
                    same = ok if same is None else (same and ok)
                loc += n_here
            t = torch.tensor([1 if same else 0], device=dev)
            d.dist.all_reduce(t, op=d.dist.ReduceOp.MIN)
            parity["neighbour_rows_bit_exact"] = bool(t.item() == 1)
            parity["ok"] = bool(parity["ok"] and parity["neighbour_rows_bit_exact"])
    del weights

    # ---- end to end from host token strings ("e2e") -----------------------------------------------------------------
    e2e = None
    if want_e2e:
        pipe = TokenPipeline(nat, hn, src, lang, rows_per_pass=P, first_chunk_rows=env_int("ZETT_BENCH_FIRST_CHUNK", -1) if "ZETT_BENCH_FIRST_CHUNK" in os.environ else None)

        def e2e_step():
            # host token strings -> native retokenizer -> pinned H2D -> forward into this rank's slots -> in-place all-gather
            # per super-block on the side stream -> pinned D2H of this rank's rows; all of it pipelined per pass
            return pipe.run(tokens, comm=d.comm if world > 1 else None, full=full, plan=plan, rank=rank, side=getattr(d, "side", None))

        e2e_step()
        d.sync_all()
        t0 = time.perf_counter()
        for _ in range(steps):
            e2e_step()
        d.sync_all()
        e2e_s = d.max_over_ranks(time.perf_counter() - t0) / steps
        out_pinned, sf_e2e, _ = e2e_step()
        assert np.isfinite(out_pinned.numpy()[:, : (2 if separate else 1) * D + 1]).all()
        assert np.array_equal(sf_e2e, sf_host)
        e2e = {"value": world * rows / e2e_s, "unit": "rows/s", "h2d_bytes_per_step": int(world * rows * cfg.hn_surface_maxlen * 4),
               "d2h_bytes_per_step": int(world * rows * width * 4), "ms_per_step": e2e_s * 1e3}
    nat.close()
    if world > 1:
        d.comm.unregister()
    del src, sf_dev, full
    torch.cuda.empty_cache()

    peaks = measured_peaks()
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    f_ref = cfg.flops_per_row(pruned=False)
    res = {
        "value": value, "ms_per_step": ms_per_step, "dtype": fmt[0], "e2e": e2e, "clocks": clocks, "parity": parity,
        "gpu_launches": int(launches * steps), "steps_timed": int(steps),
        "config": {
            "workload": wl["workload"] + ("" if vocab == "random" else " [target vocabulary: concatenations of hn pieces]"),
            "rows_per_gpu": rows, "total_rows": world * rows, "parallelism": "rows x%d" % world,
            "gather": {"none": "single GPU", "p2p": "zett_allgather_rows: peer copies per super-block on a side stream (copy engines, NVLink)",
                       "nccl": "zett_allgather_rows: ncclAllGather per super-block on a side stream"}.get(transport, transport),
            "hn_tokenizer": (hn_kind or wl["hn"]) + " 32k synthetic", "nonpad_length_histogram": hist, "truncated": n_trunc,
            "l2": "inputs larger than L2 (weights + per-pass activations are GBs; nothing is re-read from a warm L2 by design)",
            "gemm_impl": int(st["gemm_impl"]), "split_terms": terms, "rows_per_pass": P,
            "distinct_ids_per_step": int(distinct_ids), "distinct_id_position_pairs_per_step": int(distinct_pairs),
            "packed_positions_per_step": int(positions),
            "dense_gflop_per_row_reference": f_ref / 1e9,
            "executed_gflop_per_row": gemm_flops / rows / 1e9 if gemm_flops else None,
        },
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                     "frac": (achieved / peaks["tflops"]) if achieved else None, "traffic": None,
                     "kernel": "gemm_tcgen05_kernel", "peak_source": peaks["src"],
                     "note": "achieved = FLOPs of the GEMMs issued in one step (each product counted once although the operand "
                             "format issues several MMA terms per product) / summed CUDA-event time of those launches",
                     "gemm_ms_per_step": gemm_ms, "gemm_launches_per_step": gemm_launches,
                     "mma_issue_tflops_f16_equivalent": (achieved * fmt[2]) if achieved else None},
    }
    prof = profile_of_largest_gemm(name, fmt[1], int(st["gemm_impl"]) == 5)
    if prof:
        res["roofline"]["traffic"] = prof["traffic"]
        res["roofline"]["profile"] = prof["note"]
    return res


def slim(res):
    """An `extra` entry: the figures of a secondary workload without the long config block."""
    keep = ("rows_per_gpu", "total_rows", "workload", "hn_tokenizer", "distinct_ids_per_step", "distinct_id_position_pairs_per_step",
            "packed_positions_per_step", "executed_gflop_per_row", "split_terms", "gemm_impl")
    return {"value": res["value"], "unit": "rows/s", "ms_per_step": res["ms_per_step"], "steps": res.get("steps_timed"), "e2e": res["e2e"],
            "clocks": res["clocks"],
            "parity": res["parity"], "roofline": {k: res["roofline"][k] for k in ("achieved", "peak", "frac", "gemm_ms_per_step")},
            "config": {k: res["config"][k] for k in keep}}


def run_ours(args):
    d = Dist()
    world, rank = d.world, d.rank
    main = measure(d, args.config, args)
    extra = {}
    if not args.no_extra:
        if world == 1:
            # the other single-GPU configurations of BASELINE.json, in the same run
            for other in ("xlmr", "tinyllama", "mistral"):
                if other != args.config:
                    extra[other] = slim(measure(d, other, args, steps=max(3, args.steps), min_region_ms=800.0))
            # how much the headline leans on the de-duplication the synthetic vocabulary allows
            extra["%s_all_distinct" % args.config] = slim(measure(
                d, args.config, args, env={"ZETT_DEDUP_IDS": "0", "ZETT_DEDUP_PAIRS": "0"}, want_e2e=False, want_parity=False, steps=2))
            extra["%s_all_distinct" % args.config]["what"] = ("same workload with both de-duplications switched off (input projection per "
                                                              "position, first encoder layer per position): every position pays full price")
            extra["%s_concat_vocab_unigram_hn" % args.config] = slim(measure(
                d, args.config, args, vocab="concat", hn_kind="unigram", want_e2e=False, steps=2))
            extra["%s_concat_vocab_unigram_hn" % args.config]["what"] = (
                "target vocabulary built from concatenations of pieces of a Unigram hn tokenizer: ids spread over the whole source "
                "table (see distinct_ids_per_step), de-duplication on")
        if world == 8 and args.config == "mistral":
            # BASELINE.json configs[4]: Mistral-7B hypernet, synthetic 256k vocabulary row-sharded across 8 GPUs
            extra["mistral_256k_vocab_8gpu"] = slim(measure(d, "mistral", args, rows=32768, steps=max(3, args.steps), min_region_ms=800.0))
            extra["mistral_256k_vocab_8gpu"]["what"] = "BASELINE.json configs[4]: 262 144 rows, 32 768 per GPU, all-gather per super-block"
    if rank != 0:
        d.close()
        return
    line = {
        "metric": METRIC, "value": main["value"], "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": main["dtype"], "data": "synthetic", "config": main["config"], "clocks": main["clocks"], "e2e": main["e2e"],
        "gpu_launches": main["gpu_launches"], "roofline": main["roofline"], "parity": main["parity"],
    }
    if extra:
        line["extra"] = extra
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(args.config, 1, 1, budget_s=args.cpu_seconds)
        line["cpu_baseline"] = {"value": r["value"], "unit": "rows/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
        try:
            line["torch_eager_b200"] = torch_eager_gpu_run(args.config, d.dev)
        except Exception as e:  # noqa: BLE001
            line["torch_eager_b200"] = {"error": str(e)[:200]}
    print(json.dumps(line), flush=True)
    d.close()


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
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads of the `extra` block")
    ap.add_argument("--parity-rows", type=int, default=128)
    ap.add_argument("--cpu-seconds", type=float, default=30.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
