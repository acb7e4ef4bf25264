"""Row-sharded multi-GPU prediction: one process per GPU, contiguous row blocks, ONE all-gather at the end.

The reference shards each inference batch over the local devices with ``PositionalSharding`` and keeps the hypernet
parameters and the source-embedding table replicated (scripts/transfer.py:90-91, zett/utils.py:26).  Rows are
independent (no cross-row operation in any supported configuration), so here rank ``r`` of ``G`` predicts rows
``[r * ceil(V / G), (r + 1) * ceil(V / G))`` with no data-path communication, writing ``pred_in | pred_out | bias``
of each row side by side into one ``[rows_per_rank, n_out * D + 4]`` fp32 block (the kernels take the row stride),
and a single ``all_gather_into_tensor`` (NCCL over NVLink/NVSwitch; gloo in the CPU tests) assembles the full
matrix on every rank.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

BIAS_PAD = 4  # the bias column is padded to 4 floats so that every row of the block stays 16-byte aligned


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int, int]:
    """(lo, hi, rows_per_rank): rank's rows are [lo, hi); every rank's block holds rows_per_rank rows (tail padded)."""
    per = max(1, math.ceil(n_rows / world))
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per), per


def packed_width(n_embd: int, separate_out: bool) -> int:
    return (2 if separate_out else 1) * n_embd + BIAS_PAD


def gather_rows(block: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """The single collective of the path: [rows_per_rank, W] per rank -> [world * rows_per_rank, W] on every rank."""
    if world == 1:
        return block
    full = torch.empty((world * block.shape[0], block.shape[1]), dtype=block.dtype, device=block.device)
    dist.all_gather_into_tensor(full, block.contiguous(), group=group)
    return full


def unpack(full: torch.Tensor, n_rows: int, n_embd: int, separate_out: bool):
    d = n_embd
    pred_in = full[:n_rows, :d]
    pred_out = full[:n_rows, d:2 * d] if separate_out else None
    bias = full[:n_rows, (2 if separate_out else 1) * d]
    return pred_in, pred_out, bias


def predict_sharded(n_rows: int, n_embd: int, separate_out: bool, compute_block: Callable[[int, int, torch.Tensor], None],
                    device, world: Optional[int] = None, rank: Optional[int] = None, group=None):
    """Generic driver: ``compute_block(lo, hi, block)`` must fill ``block[: hi - lo]`` for rows [lo, hi)."""
    world = dist.get_world_size(group) if world is None else world
    rank = dist.get_rank(group) if rank is None else rank
    lo, hi, per = shard_bounds(n_rows, world, rank)
    block = torch.zeros((per, packed_width(n_embd, separate_out)), dtype=torch.float32, device=device)
    if hi > lo:
        compute_block(lo, hi, block)
    full = gather_rows(block, world, group)
    check = getattr(compute_block, "check", None)
    if check is not None:
        check()  # IndexError for an out-of-range id, as ZettHypernet.forward raises (the kernels clamp and only flag it)
    return unpack(full, n_rows, n_embd, separate_out)


def hypernet_block_fn(hypernet, surface_forms_dev: torch.Tensor, source_embeddings_dev: torch.Tensor, lang_index=None):
    """``compute_block`` for a ``zett_b200.ZettHypernet``: the kernels write straight into the packed block."""
    nat = hypernet.native(surface_forms_dev.device)
    cfg = hypernet.config
    d = cfg.n_embd
    separate = bool(cfg.separate_out_embeddings)
    lang = -1 if (lang_index is None or not cfg.hn_embed_lang_id) else int(lang_index)

    def compute(lo: int, hi: int, block: torch.Tensor):
        w = block.shape[1]
        sf = surface_forms_dev[lo:hi].contiguous()
        nat.forward_into(sf, source_embeddings_dev, lang, block[:, 0:], block[:, d:] if separate else None,
                         block[:, (2 if separate else 1) * d:], ld_pred=w, ld_bias=w)

    compute.check = nat.check
    return compute
