"""Row-sharded multi-GPU prediction: one process per GPU, no data-path communication, ONE collective kind (all-gather).

The reference shards each inference batch over the local devices with ``PositionalSharding`` and keeps the hypernet
parameters and the source-embedding table replicated (scripts/transfer.py:90-91, 105-111; zett/utils.py:26).  Rows are
independent (no cross-row operation in any supported configuration), so the vocabulary is cut the same way: into
SUPER-BLOCKS of ``world * rows_per_pass`` consecutive rows, of which rank ``r`` predicts the ``r``-th slice of
``rows_per_pass`` rows with its own replica of the weights.  The kernels write ``pred_in | pred_out | bias`` of a row
side by side straight into this rank's slot of the full ``[rows, n_out * D + 4]`` fp32 matrix (they take the row
stride), and one in-place all-gather per super-block -- ``zett_allgather_rows`` of libzett_b200.so, enqueued on a side
stream so that it runs under the next super-block's compute -- fills in the other ranks' slots.  Its transport is
ncclAllGather, or, once the full matrix is registered with the communicator (``NativeComm.register``), peer copies over
NVLink / NVSwitch that need no SM (the persistent GEMMs own all of them).  Only the last super-block's gather is exposed.
The CPU tests drive the same plan over gloo.
"""
from __future__ import annotations

import ctypes
import math
from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist

BIAS_PAD = 4  # the bias column is padded to 4 floats so that every row of the block stays 16-byte aligned


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int, int]:
    """(lo, hi, rows_per_rank) of ONE super-block covering all rows: rank's rows are [lo, hi); every rank's slot holds
    rows_per_rank rows (tail padded)."""
    per = max(1, math.ceil(n_rows / world))
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per), per


def shard_plan(n_rows: int, world: int, rows_per_pass: int) -> List[Tuple[int, int]]:
    """Super-blocks as (first row, rows per rank): block s covers rows [base, base + world * per); all but the last one
    have per == rows_per_pass, the last one spreads what is left evenly (its padding falls past row n_rows)."""
    plan, base = [], 0
    rows_per_pass = max(1, int(rows_per_pass))
    while base < n_rows:
        per = min(rows_per_pass, math.ceil((n_rows - base) / world))
        plan.append((base, per))
        base += world * per
    return plan


def padded_rows(n_rows: int, world: int, rows_per_pass: int) -> int:
    plan = shard_plan(n_rows, world, rows_per_pass)
    return (plan[-1][0] + world * plan[-1][1]) if plan else 0


def packed_width(n_embd: int, separate_out: bool) -> int:
    return (2 if separate_out else 1) * n_embd + BIAS_PAD


def unpack(full: torch.Tensor, n_rows: int, n_embd: int, separate_out: bool):
    d = n_embd
    pred_in = full[:n_rows, :d]
    pred_out = full[:n_rows, d:2 * d] if separate_out else None
    bias = full[:n_rows, (2 if separate_out else 1) * d]
    return pred_in, pred_out, bias


# ----------------------------------------------------------------------------------------------------------------------
# communicators: the library's NCCL binding on GPUs, torch.distributed (gloo) for the CPU tests of the host logic
# ----------------------------------------------------------------------------------------------------------------------
class NativeComm:
    """``zett_comm`` of libzett_b200.so (``zett_comm_init`` / ``zett_allgather_rows`` = ncclAllGather)."""

    def __init__(self, rank: int, world: int, unique_id: Optional[bytes]):
        from . import _lib
        self._lib = _lib
        self.lib = _lib.load()
        self.rank, self.world = int(rank), int(world)
        self.handle = ctypes.c_void_p()
        buf = ctypes.create_string_buffer(unique_id, 128) if unique_id is not None else None
        _lib.check(self.lib.zett_comm_init(self.rank, self.world, buf, ctypes.byref(self.handle)))

    @staticmethod
    def unique_id() -> bytes:
        from . import _lib
        buf = ctypes.create_string_buffer(128)
        _lib.check(_lib.load().zett_comm_unique_id(buf))
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, group=None) -> "NativeComm":
        """Bootstrap over an initialised torch.distributed group: rank 0 mints the NCCL id, everybody receives it.
        The current CUDA device of each process becomes its device in the communicator."""
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.unique_id() if rank == 0 else None]
        if world > 1:
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return cls(rank, world, box[0] if world > 1 else None)

    def allgather_rows(self, full_slab: torch.Tensor, per: int, stream: Optional[torch.cuda.Stream] = None):
        """In place: this rank's ``per`` rows already sit in its slot of ``full_slab[world * per, W]``."""
        if self.world == 1:
            return
        assert full_slab.is_contiguous() and full_slab.dtype == torch.float32 and full_slab.shape[0] == self.world * per
        width = full_slab.shape[1]
        shard_ptr = full_slab.data_ptr() + self.rank * per * width * 4
        s = (stream or torch.cuda.current_stream(full_slab.device)).cuda_stream
        self._lib.check(self.lib.zett_allgather_rows(self.handle, ctypes.c_void_p(shard_ptr), per, width,
                                                     ctypes.c_void_p(full_slab.data_ptr()), ctypes.c_void_p(s)))

    def register(self, full: torch.Tensor, group=None) -> bool:
        """Collective: from now on ``allgather_rows`` into ``full`` is peer copies (copy engines over NVLink, no kernel) instead
        of ncclAllGather.  ``full`` must have the same size on every rank.  Handles travel over torch.distributed.
        Returns False -- on EVERY rank, with nothing registered -- when any rank cannot export or map the buffers (no peer
        access, CUDA IPC not permitted in the container): the caller then simply keeps the ncclAllGather transport."""
        if self.world == 1:
            return True
        handle, off = ctypes.create_string_buffer(64), ctypes.c_int64(0)
        rc = self.lib.zett_comm_ipc_handle(ctypes.c_void_p(full.data_ptr()), handle, ctypes.byref(off))
        box = [None] * self.world
        dist.all_gather_object(box, (handle.raw, int(off.value)) if rc >= 0 else None, group=group)
        if any(b is None for b in box):
            return False
        blob = b"".join(h for h, _ in box)
        offs = (ctypes.c_int64 * self.world)(*[o for _, o in box])
        rc = self.lib.zett_comm_register(self.handle, ctypes.c_void_p(full.data_ptr()), full.numel() * full.element_size(), blob, offs)
        oks = [None] * self.world
        dist.all_gather_object(oks, rc >= 0, group=group)
        if not all(oks):
            if rc >= 0:
                torch.cuda.synchronize(full.device)
                self.lib.zett_comm_unregister(self.handle)
            dist.barrier(group=group)
            return False
        self._registered = full   # keep the tensor alive while peers map it
        return True

    def unregister(self, group=None):
        if self.world > 1 and getattr(self, "_registered", None) is not None:
            torch.cuda.synchronize(self._registered.device)
            dist.barrier(group=group)   # nobody frees its buffer while a peer may still be pushing into it
            self._lib.check(self.lib.zett_comm_unregister(self.handle))
            dist.barrier(group=group)
            self._registered = None

    def barrier(self, stream: Optional[torch.cuda.Stream] = None):
        """Completes (in stream order) when every rank has reached it: all peer copies enqueued before it have landed."""
        if self.world == 1:
            return
        s = (stream or torch.cuda.current_stream()).cuda_stream
        self._lib.check(self.lib.zett_comm_barrier(self.handle, ctypes.c_void_p(s)))

    def info(self) -> dict:
        r, w, v = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._lib.check(self.lib.zett_comm_info(self.handle, ctypes.byref(r), ctypes.byref(w), ctypes.byref(v)))
        return {"rank": r.value, "world": w.value, "nccl_version": v.value}

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.zett_comm_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class TorchComm:
    """The same contract over torch.distributed (gloo on CPU): used by the world_size-2 CPU tests."""

    def __init__(self, group=None, world: Optional[int] = None, rank: Optional[int] = None):
        self.group = group
        self.world = dist.get_world_size(group) if world is None else world
        self.rank = dist.get_rank(group) if rank is None else rank

    def allgather_rows(self, full_slab: torch.Tensor, per: int, stream=None):
        if self.world == 1:
            return
        shard = full_slab[self.rank * per:(self.rank + 1) * per].clone()
        dist.all_gather_into_tensor(full_slab, shard, group=self.group)

    def barrier(self, stream=None):
        """(torch.distributed collectives complete for every rank by themselves)"""


def gather_rows(block: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """[rows_per_rank, W] per rank -> [world * rows_per_rank, W] on every rank through torch.distributed."""
    if world == 1:
        return block
    full = torch.empty((world * block.shape[0], block.shape[1]), dtype=block.dtype, device=block.device)
    dist.all_gather_into_tensor(full, block.contiguous(), group=group)
    return full


# ----------------------------------------------------------------------------------------------------------------------
# drivers
# ----------------------------------------------------------------------------------------------------------------------
def predict_sharded(n_rows: int, n_embd: int, separate_out: bool, compute_block: Callable[[int, int, torch.Tensor], None],
                    device, world: Optional[int] = None, rank: Optional[int] = None, group=None, comm=None,
                    rows_per_pass: Optional[int] = None, full: Optional[torch.Tensor] = None):
    """Generic driver: ``compute_block(lo, hi, slot)`` must fill ``slot[: hi - lo]`` (a view into the full matrix) for
    vocabulary rows [lo, hi).  ``rows_per_pass`` None = one super-block (one gather at the very end).  On CUDA with a
    ``NativeComm`` the gather of a super-block runs on a side stream under the compute of the next one."""
    if comm is None:
        comm = TorchComm(group, world, rank)
    world, rank = comm.world, comm.rank
    plan = shard_plan(n_rows, world, rows_per_pass if rows_per_pass else max(1, math.ceil(n_rows / world)))
    width = packed_width(n_embd, separate_out)
    rows_pad = (plan[-1][0] + world * plan[-1][1]) if plan else 0
    if full is None or full.shape[0] < rows_pad:
        full = torch.zeros((rows_pad, width), dtype=torch.float32, device=device)
    on_gpu = torch.device(device).type == "cuda"
    side = torch.cuda.Stream(device=device) if (on_gpu and world > 1) else None
    for base, per in plan:
        lo = min(n_rows, base + rank * per)
        hi = min(n_rows, lo + per)
        slab = full[base:base + world * per]
        if hi > lo:
            compute_block(lo, hi, slab[rank * per:rank * per + (hi - lo)])
        if side is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(device))
            side.wait_event(ev)
            comm.allgather_rows(slab, per, stream=side)
        else:
            comm.allgather_rows(slab, per)
    if side is not None:
        torch.cuda.current_stream(device).wait_stream(side)
    comm.barrier()   # peer copies complete locally: everybody's rows have landed once every rank is past this point
    check = getattr(compute_block, "check", None)
    if check is not None:
        check()  # IndexError for an out-of-range id in ANY pass, as ZettHypernet.forward raises (the kernels clamp and flag)
    return unpack(full, n_rows, n_embd, separate_out)


def hypernet_block_fn(hypernet, surface_forms_dev: torch.Tensor, source_embeddings_dev: torch.Tensor, lang_index=None):
    """``compute_block`` for a ``zett_b200.ZettHypernet`` (or a ``NativeHypernet``): the kernels write straight into the
    rank's slot of the full matrix."""
    nat = hypernet.native(surface_forms_dev.device) if hasattr(hypernet, "native") else hypernet
    cfg = nat.cfg
    d = cfg.n_embd
    separate = bool(cfg.separate_out_embeddings)
    lang = -1 if (lang_index is None or not cfg.hn_embed_lang_id) else int(lang_index)

    def compute(lo: int, hi: int, slot: torch.Tensor):
        w = slot.shape[1]
        sf = surface_forms_dev[lo:hi].contiguous()
        nat.forward_into(sf, source_embeddings_dev, lang, slot[:, 0:], slot[:, d:] if separate else None,
                         slot[:, (2 if separate else 1) * d:], ld_pred=w, ld_bias=w)

    compute.check = nat.check
    return compute
