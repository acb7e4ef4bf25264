"""Multi-GPU parity check (run under torchrun on N GPUs of one box; tests/test_multi_gpu.py launches it): the row-sharded
prediction assembled by the library's own all-gather (zett_comm_init / zett_allgather_rows = ncclAllGather, one per
super-block, on a side stream) equals the single-GPU prediction bit for bit, on every rank, and matches the oracle on a
sample."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import hypernet_oracle as ho  # noqa: E402
import zett_synthetic as synthetic
from zett_b200 import parallel  # noqa: E402
from zett_b200.modeling_hypernet import ZettHypernet, load_weights_numpy  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    comm = parallel.NativeComm.from_torch_distributed()
    info = comm.info()
    assert info["world"] == world and info["rank"] == rank and (world == 1 or info["nccl_version"] > 0), info
    ok = True
    for name, rows, lang in (("tiny", 1003, None), ("xlmr", 4099, 3)):
        cfg = synthetic.make_config(name)
        weights = synthetic.make_weights(cfg, seed=11)
        src_np = synthetic.make_source_embeddings(cfg, seed=12)
        sf_np = synthetic.make_random_surface_forms(cfg, rows, seed=31)
        model = load_weights_numpy(ZettHypernet(cfg), weights).to(dev)
        sf, src = torch.from_numpy(sf_np).to(dev), torch.from_numpy(src_np).to(dev)
        single = model(sf, source_embeddings=src, lang_index=None if lang is None else torch.tensor(lang))
        fn = parallel.hypernet_block_fn(model, sf, src, lang)
        for rpp, transport in ((None, "nccl"), (96, "nccl"), (96, "p2p")):
            full = None
            if transport == "p2p":   # registered full matrix: the gather becomes peer copies over NVLink
                full = torch.zeros((parallel.padded_rows(rows, world, rpp), parallel.packed_width(cfg.n_embd, bool(cfg.separate_out_embeddings))),
                                   dtype=torch.float32, device=dev)
                assert comm.register(full), "peer registration failed on this box"
            sharded = parallel.predict_sharded(rows, cfg.n_embd, bool(cfg.separate_out_embeddings), fn, dev, comm=comm,
                                               rows_per_pass=rpp, full=full)
            torch.cuda.synchronize()
            if transport == "p2p":
                sharded = tuple(None if t is None else t.clone() for t in sharded)
                comm.unregister()
            for a, b in zip(single, sharded):
                if a is not None and not torch.equal(a, b.contiguous()):
                    ok = False
                    print(f"rank {rank} {name} rows_per_pass {rpp} {transport}: sharded != single, max diff {(a - b).abs().max().item():.3e}", flush=True)
        for a, b in zip(single, sharded):
            if a is None:
                assert b is None
                continue
            same = torch.equal(a, b.contiguous())
            ok &= same
            if not same:
                print(f"rank {rank} {name}: sharded != single, max diff {(a - b).abs().max().item():.3e}", flush=True)
        if rank == 0:
            want = ho.hypernet_forward(cfg, weights, sf_np[:256], src_np, lang_index=lang)
            masked = ho.fully_masked_rows(cfg, sf_np[:256])
            for g, w in zip(sharded, want):
                if w is None:
                    continue
                fro, worst = ho.rel_errors(g[:256].cpu().numpy(), w, exclude=masked)
                ok &= fro < 1e-3 and worst < 1e-3
                print(f"{name}: world {world} vs oracle fro {fro:.2e} worst {worst:.2e}", flush=True)
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if t.item() == 1 else "FAIL", "world", world, "nccl", info["nccl_version"], flush=True)
    comm.close()
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1 else 1)


if __name__ == "__main__":
    main()
