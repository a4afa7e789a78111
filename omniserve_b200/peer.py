"""Symmetric (peer-mapped) buffers for the fused tensor-parallel all-reduce (csrc/small_ops.cu: PeerCtx).

PyTorch is plumbing here: `torch.distributed._symmetric_memory` allocates a buffer on every GPU of the node and maps all
of them into every process (CUDA peer access over NVLink); the kernels of this repository do the communication
themselves with plain loads / stores and release / acquire flags in that memory.  The reference has no counterpart (no
TP).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

MAX_BLOCKS = 256  # token rows per call (one block per row); decode batches only


class PeerGroup:
    """Flag and epoch arrays shared by all PeerBuffers of one process group (one set is enough: every rank executes the
    same sequence of fused all-reduce calls, so per-block epochs stay in lock-step)."""

    def __init__(self, group, device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if not 2 <= self.world <= 8:
            raise RuntimeError("peer all-reduce supports 2..8 GPUs of one node")
        import warnings
        with warnings.catch_warnings():   # a no-op (and deprecated) on recent PyTorch, required on older ones
            warnings.simplefilter("ignore")
            try:
                symm.enable_symm_mem_for_group(self.group.group_name)
            except Exception:
                pass
        self._symm = symm
        self.device = device
        self.flags = symm.empty(MAX_BLOCKS * 8, dtype=torch.int32, device=device)
        self.flags.zero_()
        self.flags_hdl = symm.rendezvous(self.flags, self.group)
        self.epoch = torch.zeros(MAX_BLOCKS, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(self.group)     # every rank's flags are zero before anyone signals

    def buffer(self, rows: int, cols: int) -> "PeerBuffer":
        return PeerBuffer(self, rows, cols)


class PeerBuffer:
    """A [rows, cols] fp16 tensor on every rank, each mapped into all ranks.  `tensor` is the local one (the row-parallel
    GEMM writes its partial sums into it); `ctx_ref()` is the ob_peer_ctx the fused norm kernels take."""

    def __init__(self, pg: PeerGroup, rows: int, cols: int):
        if rows > MAX_BLOCKS:
            raise RuntimeError(f"peer all-reduce handles at most {MAX_BLOCKS} token rows")
        self.pg = pg
        self.tensor = pg._symm.empty((rows, cols), dtype=torch.float16, device=pg.device)
        self.tensor.zero_()
        self.hdl = pg._symm.rendezvous(self.tensor, pg.group)
        ctx = L.PeerCtx()
        bufs, flags = list(self.hdl.buffer_ptrs), list(pg.flags_hdl.buffer_ptrs)
        for p in range(pg.world):
            ctx.bufs[p] = bufs[p]
            ctx.flags[p] = flags[p]
        ctx.epoch = pg.epoch.data_ptr()
        ctx.world, ctx.rank, ctx.max_blocks = pg.world, pg.rank, MAX_BLOCKS
        self._ctx = ctx

    def ctx_ref(self):
        return C.byref(self._ctx)
