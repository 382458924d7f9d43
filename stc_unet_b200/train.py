"""Training-step plumbing around the hot path:

  * FlatParams   — every parameter becomes a view into ONE fp32 buffer (one rank-0 broadcast, one fused Adam kernel)
  * GradArena    — (ops.py) every parameter gradient is written by our kernels straight into ONE fp32 buffer
  * GradReducer  — bucketed NCCL all-reduce(AVG) of that buffer on a side stream, launched from post-accumulate
                   hooks while backward is still running (replaces MMDistributedDataParallel's reducer,
                   mmseg/apis/train.py:104-113, broadcast_buffers=False)
  * FusedAdam    — torch.optim.Adam semantics (my_config/STC-UNet.py:87: lr 1e-5, betas (0.9, 0.999)) as one
                   stc_adam_step launch over the flat buffers
  * Trainer      — zero_grad -> forward_train -> backward -> (all-reduce) -> Adam, i.e. what mmcv's
                   EpochBasedRunner + OptimizerHook do per iteration around BaseSegmentor.train_step.

SyncBN statistics are exchanged inside the BN ops themselves (ops._bn_forward_stats / _bn_backward) whenever the BN
containers are nn.SyncBatchNorm and torch.distributed is initialised.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from . import ops
from . import peer as peer_mod
from ._lib import lib, stream_ptr


def _dist_on() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class FlatParams:
    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        dev = self.params[0].device
        total, self.offsets = 0, []
        for p in self.params:
            self.offsets.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        for p, off in zip(self.params, self.offsets):
            view = self.flat[off:off + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
        self.total = total

    def broadcast(self, src: int = 0, buffers=()):
        """Rank `src`'s parameters - and, once at start-up, the module buffers (BN running statistics), as DDP does at construction -
        to every rank.  Afterwards the buffers stay identical by construction (every rank applies the same global SyncBN statistics;
        the reference passes broadcast_buffers=False, mmseg/apis/train.py:112)."""
        if _dist_on():
            dist.broadcast(self.flat, src)
            for b in buffers:
                dist.broadcast(b, src)


def plan_buckets(spans, total: int, cap_elems: int):
    """spans: [(offset, numel)] of the parameters in arena (= forward) order.  Returns (buckets, owner): buckets are
    contiguous [start, end) ranges taken from the END of the arena (backward produces the last parameters first), each
    at least `cap_elems` long (except possibly the last one), with the number of parameters inside; owner[i] is the
    bucket index of parameter i."""
    buckets, owner = [], [0] * len(spans)
    cur_end, cur_start, cnt = total, total, 0
    for i in range(len(spans) - 1, -1, -1):
        cur_start = spans[i][0]
        owner[i] = len(buckets)
        cnt += 1
        if cur_end - cur_start >= cap_elems:
            buckets.append((cur_start, cur_end, cnt))
            cur_end, cnt = cur_start, 0
    if cnt:
        buckets.append((cur_start, cur_end, cnt))
    return buckets, owner


class GradReducer:
    """Buckets are contiguous ranges of the arena taken from its END (backward produces the last parameters
    first); a bucket's all-reduce starts on the comm stream as soon as all of its gradients have been written."""

    def __init__(self, arena: ops.GradArena, bucket_mb: float = 25.0, peer=None):
        self.arena = arena
        self.peer = peer          # peer.PeerExchange: the buckets are reduced by our NVLink peer-memory kernel instead of NCCL
        self.comm = torch.cuda.Stream()
        cap = int(bucket_mb * (1 << 20) // 4)
        spans = [arena.offsets[id(p)] for p in arena.params]
        self.buckets, owner = plan_buckets(spans, arena.total, cap)   # (start, end, n_params)
        self.bucket_of: Dict[int, int] = {id(p): owner[i] for i, p in enumerate(arena.params)}
        self.pending = [0] * len(self.buckets)
        self.handles = []
        self.enabled = _dist_on()
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in arena.params]

    def reset(self):
        self.pending = [b[2] for b in self.buckets]
        self.handles = []

    def _on_grad(self, p):
        g = p.grad
        if g is not None:      # a gradient that is not the arena slice itself (foreign op, shared weight summed by autograd): copy it in
            v = self.arena.view(p)
            if g.data_ptr() != v.data_ptr():
                v.copy_(g)
        if not self.enabled:
            return
        b = self.bucket_of[id(p)]
        self.pending[b] -= 1
        if self.pending[b] == 0:
            self._launch(b)

    def _launch(self, b):
        start, end, _ = self.buckets[b]
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.comm.wait_event(ev)
        with torch.cuda.stream(self.comm):
            if self.peer is not None:
                self.peer.allreduce_arena_(start, end, average=True)
            else:
                self.handles.append(dist.all_reduce(self.arena.flat[start:end], op=dist.ReduceOp.AVG, async_op=True))
        self.pending[b] = -1      # launched

    def finish(self):
        if not self.enabled:
            return
        # A bucket whose parameters did not all receive a gradient this step (an unused branch) never reached zero: reduce it now.
        # Every rank runs the same graph, so every rank flushes the same buckets in the same order and the exchanges still pair up;
        # the missing gradients are the zeros the step's memset left in the arena (DDP would raise for unused parameters instead).
        for b in range(len(self.buckets)):
            if self.pending[b] > 0:
                self._launch(b)
        for h in self.handles:
            h.wait()
        torch.cuda.current_stream().wait_stream(self.comm)
        self.handles = []


class FusedAdam:
    """Adam over the flat parameter / gradient buffers (one kernel).  Exposes `param_groups` so LR schedulers
    written against torch.optim (mmcv's PolyLrUpdaterHook sets group['lr']) keep working."""

    def __init__(self, flat_params: FlatParams, arena: ops.GradArena, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        assert flat_params.total == arena.total
        self.fp, self.arena = flat_params, arena
        self.param_groups = [dict(params=flat_params.params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, initial_lr=lr)]
        self.m = torch.zeros_like(flat_params.flat)
        self.v = torch.zeros_like(flat_params.flat)
        self.t = 0
        # device-resident copies of (lr, beta^t) and the step counter for the graph-capturable update
        self.dyn = torch.zeros(3, dtype=torch.float32, device=flat_params.flat.device)
        self.step_count_dev = torch.zeros(1, dtype=torch.int32, device=flat_params.flat.device)
        self._lr_on_dev = None
        self._t_on_dev = None

    def zero_grad(self, set_to_none: bool = True):
        for p in self.fp.params:
            p.grad = None

    def step(self):
        g = self.param_groups[0]
        self.t += 1
        lib.call("stc_adam_step", self.fp.flat, self.arena.flat, self.m, self.v, self.fp.total, float(g["lr"]),
                 float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), self.t, stream_ptr())

    def sync_device_state(self):
        """Host -> device: step counter and learning rate (call outside a captured region, e.g. before each graph replay)."""
        lr = float(self.param_groups[0]["lr"])
        if self._lr_on_dev != lr:
            self.dyn[0:1].fill_(lr)
            self._lr_on_dev = lr
        if self._t_on_dev != self.t:        # eager steps ran in between: re-seed the device counter
            self.step_count_dev.fill_(self.t)
            self._t_on_dev = self.t

    def step_dev(self):
        """The same update with lr / step read from device memory (capturable); sync_device_state() must have run."""
        g = self.param_groups[0]
        lib.call("stc_adam_step_dev", self.fp.flat, self.arena.flat, self.m, self.v, self.fp.total, self.dyn, self.step_count_dev,
                 float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), stream_ptr())

    def state_dict(self):
        return dict(step=self.t, exp_avg=self.m, exp_avg_sq=self.v, param_groups=[{k: v for k, v in self.param_groups[0].items() if k != "params"}])

    def load_state_dict(self, sd):
        self.t = int(sd["step"])
        self.m.copy_(sd["exp_avg"]); self.v.copy_(sd["exp_avg_sq"])
        self.param_groups[0].update(sd["param_groups"][0])


class Trainer:
    def __init__(self, segmentor: torch.nn.Module, lr=1e-5, betas=(0.9, 0.999), bucket_mb=25.0):
        self.model = segmentor
        params = [p for p in segmentor.parameters() if p.requires_grad]
        self.flat = FlatParams(params)
        self.flat.broadcast(0, buffers=[b for b in segmentor.buffers() if b.is_cuda])
        # N > 1: SyncBN statistics and gradient buckets travel through NVLink peer memory with our own kernels (no NCCL inside the step)
        self.peer = peer_mod.agree(peer_mod.try_create(self.flat.flat.device), self.flat.flat.device) if _dist_on() else None
        self.arena = ops.GradArena(self.flat.params, alloc=self.peer.alloc_arena if self.peer is not None else None)
        self.reducer = GradReducer(self.arena, bucket_mb, peer=self.peer)
        self.optim = FusedAdam(self.flat, self.arena, lr=lr, betas=betas)
        self.cache = ops.StepCache()
        self.steps_done = 0
        self._graph = None
        if self.peer is not None:      # ranks leave the constructor together: the first exchange pairs up within the kernels' time-out
            torch.cuda.synchronize()
            dist.barrier()

    def step(self, img: torch.Tensor, gt_semantic_seg: torch.Tensor, _device_adam: bool = False):
        """One training iteration; returns the (device) log-var tensors, no host sync."""
        if self.peer is not None:
            self.peer.check()                # a peer wait that timed out during an earlier step invalidated it: raise here, on the host
        ops.set_grad_arena(self.arena)
        ops.set_step_cache(self.cache)
        ops.set_peer_exchange(self.peer)
        try:
            self.optim.zero_grad()
            self.arena._claimed.clear()      # (a backward pass that raised may have left its bookkeeping behind)
            self.arena.flat.zero_()          # one memset: kernels that accumulate (+=) into their gradient need zeros
            self.arena.prezeroed = True
            self.cache.begin_step()          # one batched weight-pack launch + one workspace memset (from step 2 on)
            self.reducer.reset()
            out = self.model.train_step(dict(img=img, img_metas=None, gt_semantic_seg=gt_semantic_seg))
            out["loss"].backward()
            self.reducer.finish()
            if _device_adam:
                self.optim.step_dev()
            else:
                self.optim.step()
        finally:
            ops.set_grad_arena(None)
            ops.set_step_cache(None)
            ops.set_peer_exchange(None)
            self.arena.prezeroed = False
        self.steps_done += 1
        ops.invalidate_weight_caches()      # the fused Adam moved the weights through raw pointers
        # detached: a caller holding on to the log vars must not keep this step's autograd graph (and its AccumulateGrad nodes, which
        # remember the stream they were created on) alive into a later CUDA-graph capture on another stream
        return type(out["log_vars"])((k, v.detach()) for k, v in out["log_vars"].items())

    def resync(self):
        """Call after long rank-asymmetric host work (rank-0 checkpointing / evaluation between epochs) before stepping again: the ranks
        re-enter the step together instead of relying on the exchange kernels' wall-clock bound (STC_PEER_TIMEOUT_MS, default 10 min)."""
        if _dist_on():
            torch.cuda.synchronize()
            dist.barrier()
        if self.peer is not None:
            self.peer.check()

    # ---- whole-step CUDA graph: ~1000 launches replayed as one graph (no tracing compiler involved: the kernels are ours) ----
    def capture(self, img: torch.Tensor, gt_semantic_seg: torch.Tensor):
        """Captures fwd + loss + bwd + Adam for inputs of this shape/dtype.  Runs the eager warm-up steps the StepCache needs
        first (they are real training steps).  Afterwards step_graph() replays it on new data."""
        if _dist_on() and self.peer is None:
            # tried twice at N=2 (NCCL bucket all-reduces + SyncBN exchanges inside the capture; global and thread-local capture
            # error modes): the run hangs, so the multi-GPU step stays on the eager path (+3.9 ms/step of launch gaps at N=2)
            raise RuntimeError("Trainer.capture: the captured step is single-GPU only")
        while self.steps_done < 3:                      # StepCache: record, finalize, replay
            self.step(img, gt_semantic_seg)
        self._static_img = img.clone()
        self._static_gt = gt_semantic_seg.clone()
        opt = self.optim
        opt.sync_device_state()                         # continue the eager step count on the device
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                   # one eager pass on the device-state Adam so every lazy init is done
            self.step(self._static_img, self._static_gt, _device_adam=True)
        opt.t += 1
        opt._t_on_dev = opt.t
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        multi = _dist_on()
        if multi:      # every rank enters the capture from the same state (the peer kernels pair up by ticket number)
            dist.barrier()
            torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        # thread-local capture mode with a live process group: the NCCL watchdog thread's event queries must not invalidate the capture
        with torch.cuda.graph(self._graph, capture_error_mode="thread_local" if multi else "global"):
            self._static_out = self.step(self._static_img, self._static_gt, _device_adam=True)
        return self

    def step_graph(self, img: torch.Tensor, gt_semantic_seg: torch.Tensor):
        """Replays the captured step on new data (host or device tensors; copied into the static input buffers)."""
        if self._graph is None:
            raise RuntimeError("Trainer.step_graph: call capture() first")
        if self.peer is not None:
            self.peer.check()
        if tuple(img.shape) != tuple(self._static_img.shape) or img.dtype != self._static_img.dtype:
            raise RuntimeError(f"Trainer.step_graph: captured for {tuple(self._static_img.shape)} {self._static_img.dtype}, got "
                               f"{tuple(img.shape)} {img.dtype}")
        self._static_img.copy_(img, non_blocking=True)
        self._static_gt.copy_(gt_semantic_seg, non_blocking=True)
        self.optim.sync_device_state()
        self._graph.replay()
        self.optim.t += 1
        self.optim._t_on_dev = self.optim.t
        self.steps_done += 1
        ops.invalidate_weight_caches()
        return self._static_out
