"""NVLink peer-memory exchanges for the data-parallel step: SyncBN statistics and the gradient all-reduce as OUR kernels
over symmetric buffers (csrc/peer.cu) instead of NCCL calls, which makes the multi-GPU step capturable as one CUDA graph.

torch.distributed._symmetric_memory only hands out the mapped peer pointers (plumbing); the rank-0 parameter broadcast at start-up and
the lazy log-var reduction stay on NCCL (they are outside the step)."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch
import torch.distributed as dist

from ._lib import lib, stream_ptr

SMALL_N = 4096          # fp64 values per small exchange (SyncBN: 2 * C <= 2048)
ARENA_CTAS = 32         # CTAs of the gradient all-reduce kernel (it overlaps with backward: keep most SMs for the convolutions)


class PeerExchange:
    def __init__(self, device: torch.device, group=None, ctas: int = ARENA_CTAS):
        import torch.distributed._symmetric_memory as symm
        self._symm = symm
        self.group = group if group is not None else dist.group.WORLD
        self.group_name = self.group.group_name
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > 16:
            raise RuntimeError("PeerExchange supports up to 16 ranks on one NVLink domain")
        self.device, self.ctas = device, int(ctas)
        nbytes = lib.raw("stc_peer_ctrl_bytes")(SMALL_N, self.ctas)
        self.ctrl = symm.empty((nbytes + 7) // 8, dtype=torch.int64, device=device)
        self.ctrl.zero_()
        hdl = symm.rendezvous(self.ctrl, self.group_name)
        self._ctrl_hdl = hdl
        self.ctrl_ptrs = (ctypes.c_ulonglong * self.world)(*[int(p) for p in hdl.buffer_ptrs])
        self.seq_small = torch.zeros(1, dtype=torch.int64, device=device)
        self.seq_arena = torch.zeros(self.ctas, dtype=torch.int64, device=device)
        self.arena = None
        self.arena_ptrs = None
        # cross-rank waits are bounded in wall-clock time (default 10 min, STC_PEER_TIMEOUT_MS); a wait that runs out raises this
        # host-visible flag instead of trapping, and check() turns it into an exception on the host
        self.err = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.timeout_ms = int(os.environ.get("STC_PEER_TIMEOUT_MS", "600000"))
        lib.call("stc_peer_configure", self.timeout_ms, self.err)
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)          # every rank's control block is zeroed before anybody's first ticket can arrive

    def check(self):
        """Raises if a cross-rank wait timed out since the last check (reads pinned host memory: no synchronisation)."""
        if int(self.err[0]) != 0:
            raise RuntimeError(f"stc_unet_b200 peer exchange: a peer did not arrive within {self.timeout_ms} ms (rank {self.rank} of {self.world}); "
                               "the step's results are invalid. Ranks must issue the same sequence of exchanges; barrier before re-entering "
                               "the step after long rank-asymmetric work, or raise STC_PEER_TIMEOUT_MS.")

    def alloc_arena(self, numel: int) -> torch.Tensor:
        """fp32 gradient arena in symmetric memory (zeroed); one per PeerExchange."""
        if self.arena is not None:
            raise RuntimeError("PeerExchange.alloc_arena: already allocated")
        self.arena = self._symm.empty(int(numel), dtype=torch.float32, device=self.device)
        self.arena.zero_()
        hdl = self._symm.rendezvous(self.arena, self.group_name)
        self._arena_hdl = hdl
        self.arena_ptrs = (ctypes.c_ulonglong * self.world)(*[int(p) for p in hdl.buffer_ptrs])
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        return self.arena

    def allreduce_small_(self, t: torch.Tensor) -> torch.Tensor:
        """In-place SUM over the ranks of a small fp64 tensor (identical result on every rank)."""
        if t.dtype != torch.float64 or not t.is_contiguous() or t.numel() > SMALL_N:
            raise RuntimeError("PeerExchange.allreduce_small_: contiguous float64 tensor of at most %d values expected" % SMALL_N)
        lib.call("stc_peer_allreduce_small_f64", ctypes.addressof(self.ctrl_ptrs), self.rank, self.world, SMALL_N, t, t, t.numel(), self.seq_small,
                 stream_ptr())
        return t

    def allreduce_arena_(self, start: int, end: int, average: bool = True):
        """In-place all-reduce of arena[start:end) on the current stream."""
        if self.arena is None:
            raise RuntimeError("PeerExchange.allreduce_arena_: call alloc_arena first")
        lib.call("stc_peer_allreduce_arena_f32", ctypes.addressof(self.arena_ptrs), ctypes.addressof(self.ctrl_ptrs), self.rank, self.world, SMALL_N,
                 int(start), int(end - start), (1.0 / self.world) if average else 1.0, self.seq_arena, self.ctas, stream_ptr())


def peer_exchange_wanted() -> bool:
    return os.environ.get("STC_PEER", "1") != "0"


def try_create(device: torch.device, group=None) -> Optional[PeerExchange]:
    """PeerExchange when torch.distributed is up with more than one rank and symmetric memory works on this box; None otherwise (the
    caller then uses NCCL for the same exchanges)."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1 and peer_exchange_wanted()):
        return None
    try:
        return PeerExchange(device, group)
    except Exception as e:  # noqa: BLE001 - any failure of the symmetric-memory rendezvous means "not available here"
        import sys
        print(f"stc_unet_b200: peer-memory exchange unavailable ({type(e).__name__}: {e}); using NCCL", file=sys.stderr)
        return None


def agree(px, device: torch.device, group=None):
    """All ranks use the peer-memory exchanges or none does: a rank whose rendezvous failed would otherwise issue NCCL calls that its
    peers never match.  One small all-reduce at start-up; returns `px` when every rank has one, else None."""
    flag = torch.tensor([1 if px is not None else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return px if int(flag) == 1 else None
