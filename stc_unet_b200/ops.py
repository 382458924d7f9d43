"""autograd.Function wrappers over the C ABI (libstc_b200.so).

PyTorch is plumbing here: it owns the device memory, the autograd tape and the streams; every
arithmetic step is one of our CUDA kernels, called with raw device pointers.  Activations are
NHWC tensors of shape (N, H, W, C) in `float32` or `bfloat16`; parameters are fp32 in the
reference's layouts (Conv2d OIHW, Linear (out,in)).
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Optional, Sequence

import torch
import torch.distributed as dist
from torch.autograd import Function

from . import _lib
from ._lib import GemmDesc, dtype_code, lib, stream_ptr

# ---------------------------------------------------------------------------------------------
# global knobs
# ---------------------------------------------------------------------------------------------
class _Config:
    engine = _lib.ENGINE_AUTO       # dense engine selection passed to stc_conv_* / stc_gemm
    fold_linear_pairs = True        # bf16 + tcgen05: q/k/v + in_proj and fc1 + fc2 of TransformerLayer run as folded GEMMs
    ksa_lazy_df = True              # KSA branch gradients are consumed implicitly by the BN backward kernels (no df tensors)
    chain_fanout_grads = True       # KSA: the input's four gradients are summed in the branch dgrads' epilogues (no add_n pass)
    # attention: softmax / its backward inside the QK^T / dO V^T products (stc_gemm_softmax*: two sweeps, no score tensor).  OPT-IN: correct to
    # 3e-5, but measured SLOWER than product + separate pass (L = 4096: 0.77 -> 1.29 ms forward, 0.90 -> 2.03 ms backward): the second sweep
    # doubles the MMA + accumulator read-out, which is what bounds a K = 256 product, and four epilogue warps do all the exponentials
    fuse_attention_softmax = os.environ.get("STC_ATTN_FUSED", "0") == "1"
    fold_eval_bn = True             # inference: eval-mode BN folded into the conv weights, activation in the conv epilogue
    widen_narrow_convs = True       # bf16: 16 / 32-channel layers are zero-padded to the tcgen05 kernels' channel granularity
    # Inference with FROZEN weights (a deployed checkpoint): keep the folded / packed bf16 operands between forwards instead of rebuilding
    # them every call.  Opt-in, because nothing can observe a raw-pointer update of a parameter (our fused Adam, load_checkpoint's in-place
    # copy through .data): call ops.invalidate_weight_caches() after changing weights, or leave this off.
    cache_eval_weights = False


config = _Config()


def _sync_world(bn) -> int:
    """World size over which BN statistics are exchanged (1 = plain BatchNorm)."""
    if bn.sync and dist.is_available() and dist.is_initialized():
        return dist.get_world_size(bn.group)
    return 1


_WS = {}


def _workspace(device, nbytes: int) -> torch.Tensor:
    key = (device.type, device.index)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf


def _chk(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("stc_unet_b200 ops need CUDA tensors (B200); there is no CPU fallback")
    lib.ensure_device(t.device.index if t.device.index is not None else torch.cuda.current_device())
    return t if t.is_contiguous() else t.contiguous()


# ---------------------------------------------------------------------------------------------
# flat gradient arena: parameter gradients are written straight into one flat fp32 buffer
# (single bucketed NCCL all-reduce + single fused Adam kernel, no per-tensor copies)
# ---------------------------------------------------------------------------------------------
class GradArena:
    def __init__(self, params: Sequence[torch.nn.Parameter], alloc=None):
        self.params = [p for p in params if p.requires_grad]
        total = 0
        self.offsets = {}
        for p in self.params:
            self.offsets[id(p)] = (total, p.numel())
            total += (p.numel() + 3) // 4 * 4  # keep every slice 16-byte aligned
        dev = self.params[0].device
        # `alloc` lets the trainer place the arena in NVLink symmetric memory (peer.PeerExchange.alloc_arena)
        self.flat = alloc(total) if alloc is not None else torch.zeros(total, dtype=torch.float32, device=dev)
        self.total = total
        self.prezeroed = False     # Trainer zeroes the whole arena once per step (one memset) and sets this
        self._claimed = set()      # parameters whose arena slice has been handed out during the CURRENT backward pass

    def claim(self, p: torch.nn.Parameter) -> bool:
        """True the first time `p` asks for its slice during one backward pass.  The set is emptied by a callback the autograd engine
        runs when that pass ends, so the bookkeeping needs no cooperation from the training loop."""
        if not self._claimed:
            try:
                torch.autograd.Variable._execution_engine.queue_callback(self._claimed.clear)
            except RuntimeError:      # not inside a backward pass (a Function.backward called by hand): nothing to track
                return True
        if id(p) in self._claimed:
            return False
        self._claimed.add(id(p))
        return True

    def view(self, p: torch.nn.Parameter) -> Optional[torch.Tensor]:
        ent = self.offsets.get(id(p))
        if ent is None:
            return None
        off, n = ent
        return self.flat[off:off + n].view(p.shape)


_ARENA: Optional[GradArena] = None
_PEER = None    # peer.PeerExchange: SyncBN statistics go through NVLink peer memory (our kernel) instead of NCCL


def set_peer_exchange(px):
    global _PEER
    _PEER = px


def _allreduce_stats(sums: torch.Tensor, bn) -> None:
    """SUM of the BN statistic vector over the ranks of the SyncBN group (C1 / C2)."""
    if _PEER is not None and (bn.group is None or bn.group is _PEER.group):
        _PEER.allreduce_small_(sums)
    else:
        dist.all_reduce(sums, group=bn.group)


def set_grad_arena(arena: Optional[GradArena]):
    global _ARENA
    _ARENA = arena


def _grad_buf(param, shape, device, zero=False) -> torch.Tensor:
    """fp32 buffer that will become param.grad: the parameter's slice of the arena when one is active.

    The slice is handed out only when nothing else lives in it: if param.grad already IS that slice (a second backward without
    zero_grad: gradient accumulation, mmcv's GradientCumulativeOptimizerHook) or the parameter was already served during this
    backward pass (a weight used twice in one graph), the kernel writes into a temporary instead and autograd ADDS it to the
    accumulated gradient - overwriting the slice would turn g1 + g2 into 2 * g2."""
    if _ARENA is not None and param is not None:
        v = _ARENA.view(param)
        if v is not None:
            g = param.grad
            aliased = g is not None and g.data_ptr() == v.data_ptr()
            if _ARENA.claim(param) and not aliased:
                if zero and not _ARENA.prezeroed:
                    v.zero_()
                return v.view(shape)
    return (torch.zeros if zero else torch.empty)(shape, dtype=torch.float32, device=device)


# ---------------------------------------------------------------------------------------------
# optional per-launch profiler (bench.py): CUDA events on the launching stream around every dense call,
# with the call's algorithmic FLOPs, plus a count of every kernel-launching C-ABI call
# ---------------------------------------------------------------------------------------------
class LaunchProfiler:
    def __init__(self, time_dense: bool = False, time_all: bool = False):
        self.time_dense = time_dense
        self.time_all = time_all  # CUDA events around EVERY C-ABI call (tools/step_profile.py)
        self.records = []       # (kind, engine, flops, start_event, end_event)
        self.all_records = []   # (entry point, bytes of the tensor arguments, start_event, end_event)
        self.all_sigs = []      # the small integer arguments of the same calls (shape signature; tools/step_profile.py --dense)
        self.calls = 0

    def dense(self, kind, flops, fn):
        if self.time_dense:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            self.records.append((kind, lib.raw("stc_dense_last_engine")(), flops, s, e))
        else:
            fn()

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for kind, engine, flops, s, e in self.records:
            d = out.setdefault((kind, engine), dict(launches=0, flops=0.0, ms=0.0))
            d["launches"] += 1
            d["flops"] += flops
            d["ms"] += s.elapsed_time(e)
        return out


_PROF: Optional[LaunchProfiler] = None


def set_profiler(p: Optional[LaunchProfiler]):
    global _PROF
    _PROF = p
    _orig = _lib._Lib.call
    if p is not None:
        def counting_call(self_, name, *args, _o=_orig):
            p.calls += 1
            if not p.time_all:
                return _o(self_, name, *args)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            _o(self_, name, *args)
            e.record()
            p.all_records.append((name, sum(a.numel() * a.element_size() for a in args if isinstance(a, torch.Tensor)), s, e))
            p.all_sigs.append(tuple(a for a in args if isinstance(a, int) and not isinstance(a, bool) and abs(a) < (1 << 20)))
        lib.call = counting_call.__get__(lib, type(lib))
    elif "call" in lib.__dict__:
        del lib.__dict__["call"]


def _dense(kind, flops, fn):
    if _PROF is not None:
        _PROF.dense(kind, flops, fn)
    else:
        fn()


# ---------------------------------------------------------------------------------------------
# dense helpers
# ---------------------------------------------------------------------------------------------
class StepCache:
    """Record / replay of a training step's weight packs and wgrad workspaces (used by train.Trainer).

    The first step records which (weight, orientation) packs and which wgrad workspaces the model asks for; from the second
    step on ONE stc_pack_conv_weights_batched launch refreshes every packed copy right after the optimizer step and ONE memset
    zeroes all wgrad workspaces, instead of ~200 small pack launches and ~100 fills per step."""

    def __init__(self):
        self.mode = "off"            # off | record | replay
        self.packs = {}              # key -> (dst_off, shape)
        self.order = []              # keys in first-use order
        self.ws_sizes, self.ws_cursor = [], 0
        self.arena = self.table = self.prefix = self.ws_arena = None
        self.dtype = None
        self.total = 0

    def begin_step(self):
        if self.mode == "off":
            self.mode = "record"
        elif self.mode == "record":
            self._finalize()
            self.mode = "replay"
        if self.mode == "replay":
            lib.call("stc_pack_conv_weights_batched", self.table, self.prefix, len(self.order), self.arena, self.total,
                     dtype_code(self.dtype), stream_ptr())
            self.ws_arena.zero_()
            self.ws_cursor = 0

    def _finalize(self):
        rows, prefix, off = [], [0], 0
        dev = None
        for key in self.order:
            ptr, shape, mode, dt, inner_pad, n, dev = self.packs[key]
            Cout, Cin, R, S = shape
            rows.append([ptr, off, Cout, Cin, R, S, inner_pad, mode])
            self.packs[key] = (off, n)
            off += (n + 7) // 8 * 8            # keep every packed tensor 16-byte aligned (TMA base alignment)
            assert mode == 2 or inner_pad == (Cout if mode == 1 else Cin)
            prefix.append(prefix[-1] + (n if mode == 2 else Cout * Cin))   # work items: (co,ci) pairs, or elements for im2col
            self.dtype = dt
        self.total = prefix[-1]
        # dst offsets are padded, the work-item index space (prefix) is dense
        self.arena = torch.empty(max(off, 8), dtype=self.dtype, device=dev)
        self.table = torch.tensor(rows, dtype=torch.int64, device=dev)
        self.prefix = torch.tensor(prefix, dtype=torch.int64, device=dev)
        self.ws_off = [0]
        for n in self.ws_sizes:
            self.ws_off.append(self.ws_off[-1] + (n + 3) // 4 * 4)
        self.ws_arena = torch.empty(max(self.ws_off[-1], 4), dtype=torch.float32, device=dev)

    def pack(self, w, dtype, mode, inner_pad, shape_out):
        key = (w.data_ptr(), tuple(w.shape), mode, dtype, inner_pad)
        if self.mode == "replay":
            ent = self.packs.get(key)
            if ent is not None:
                off, n = ent
                return self.arena[off:off + n].view(shape_out)
            return None
        if self.mode == "record" and key not in self.packs:
            n = 1
            for d in shape_out:
                n *= d
            self.packs[key] = (w.data_ptr(), tuple(w.shape), mode, dtype, inner_pad, n, w.device)
            self.order.append(key)
        return None

    def workspace(self, n, device):
        if self.mode == "replay" and self.ws_cursor < len(self.ws_sizes) and self.ws_sizes[self.ws_cursor] == n:
            off = self.ws_off[self.ws_cursor]
            self.ws_cursor += 1
            return self.ws_arena[off:off + n]
        if self.mode == "record":
            self.ws_sizes.append(n)
        return torch.zeros(n, dtype=torch.float32, device=device)


_STEP_CACHE: Optional[StepCache] = None


def set_step_cache(c: Optional[StepCache]):
    global _STEP_CACHE
    _STEP_CACHE = c


_INFER_PACK_CACHE: dict = {}


def invalidate_weight_caches():
    """Drops the inference-time operand caches (config.cache_eval_weights)."""
    _INFER_PACK_CACHE.clear()
    _EVAL_FOLD_CACHE.clear()


def pack_weight(w: torch.Tensor, dtype: torch.dtype, transpose_flip: bool = False, im2col_pad: int = 0, cache: bool = True) -> torch.Tensor:
    """Conv2d.weight (Cout,Cin,R,S) fp32 -> packed `dtype` operand: [R*S][Cout][Cin] (fprop), [R*S][Cin][Cout] with flipped
    taps (dgrad, transpose_flip) or [1][Cout][im2col_pad] with k = tap*Cin + ci (im2col)."""
    Cout, Cin, R, S = w.shape
    if im2col_pad:
        mode, inner, shape_out = 2, im2col_pad, (1, Cout, im2col_pad)
    else:
        mode = int(transpose_flip)
        inner, outer = (Cout, Cin) if transpose_flip else (Cin, Cout)
        shape_out = (R * S, outer, inner)
    if _STEP_CACHE is not None and cache:   # cache=False: weights derived inside the step (DeconvModule) are not stable storage
        hit = _STEP_CACHE.pack(w, dtype, mode, inner, shape_out)
        if hit is not None:
            return hit
    infer_key = None
    if cache and config.cache_eval_weights and not torch.is_grad_enabled():   # opt-in: the caller promises frozen weights (see _Config)
        infer_key = (w.data_ptr(), w._version, tuple(w.shape), dtype, mode, inner)
        hit = _INFER_PACK_CACHE.get(infer_key)
        if hit is not None:
            return hit
    out = torch.empty(shape_out, dtype=dtype, device=w.device)
    lib.call("stc_pack_conv_weight", w, out, Cout, Cin, R, S, inner, mode, dtype_code(dtype), stream_ptr())
    if infer_key is not None:
        if len(_INFER_PACK_CACHE) >= 2048:
            _INFER_PACK_CACHE.clear()
        _INFER_PACK_CACHE[infer_key] = out
    return out


def _wgrad_ws(n, device):
    if _STEP_CACHE is not None:
        return _STEP_CACHE.workspace(n, device)
    return torch.zeros(n, dtype=torch.float32, device=device)


def _resize_c(x: torch.Tensor, C: int) -> torch.Tensor:
    """(..., Cx) -> (..., C): zero-padded or truncated along the channel dimension (stc_resize_channels)."""
    Cx = x.shape[-1]
    if Cx == C:
        return x
    y = torch.empty((*x.shape[:-1], C), dtype=x.dtype, device=x.device)
    lib.call("stc_resize_channels", x, y, x.numel() // Cx, Cx, C, dtype_code(x.dtype), stream_ptr())
    return y


def _widen(Cin: int, Cout: int, dtype, wgrad: bool, macs: float = 0.0):
    """Channel counts the tcgen05 kernels take (K chunks of 64 input channels; 32-multiples of output channels, 64 for wgrad) when a narrow
    bf16 layer (>= 8 channels, at most 4x padding, at least 1 GMAC - the tiny CoordAtt convs stay on the SIMT engine) would otherwise
    run on the SIMT engine; None if no widening applies."""
    if dtype != torch.bfloat16 or config.engine == _lib.ENGINE_SIMT or not config.widen_narrow_convs or macs < 1e9:
        return None
    co_q = 64 if wgrad else 32
    Cip, Cop = (Cin + 63) // 64 * 64, (Cout + co_q - 1) // co_q * co_q
    if (Cip == Cin and Cop == Cout) or Cin % 8 or Cout % 8 or Cip > 4 * Cin or Cop > 4 * Cout:
        return None
    return Cip, Cop


def conv_fprop(x, wp, bias, residual, Cout: int, R: int, S: int, act: int = 0) -> torch.Tensor:
    N, H, W, Cin = x.shape
    wide = _widen(Cin, Cout, x.dtype, False, float(N) * H * W * Cin * Cout * R * S) if residual is None else None
    if wide is not None:     # zero-padded channels change nothing in the sums; the padded outputs are dropped again
        Cip, Cop = wide
        wpp = torch.zeros((R * S, Cop, Cip), dtype=wp.dtype, device=wp.device)
        wpp[:, :Cout, :Cin] = wp.view(R * S, Cout, Cin)
        bp = None
        if bias is not None:
            bp = torch.zeros(Cop, dtype=torch.float32, device=x.device)
            bp[:Cout] = bias
        return _resize_c(conv_fprop(_resize_c(x, Cip), wpp, bp, None, Cop, R, S, act), Cout)
    y = torch.empty((N, H, W, Cout), dtype=x.dtype, device=x.device)
    _dense("conv_fprop", 2.0 * N * H * W * Cin * Cout * R * S,
           lambda: lib.call("stc_conv_fprop", x, wp, bias, residual, y, N, H, W, Cin, Cout, R, S, act, dtype_code(x.dtype),
                            config.engine, stream_ptr()))
    return y


def conv_fprop_bnstats(x, wp, bias, Cout: int, R: int, S: int):
    """y = conv(x) + bias and [sum y | sum y^2] over all pixels (fp64, 2*Cout) of the stored outputs (stc_conv_fprop_bnstats)."""
    N, H, W, Cin = x.shape
    y = torch.empty((N, H, W, Cout), dtype=x.dtype, device=x.device)
    sums = torch.empty(2 * Cout, dtype=torch.float64, device=x.device)
    ws = _workspace(x.device, lib.raw("stc_bn_ws_bytes")(N * H * W, Cout))
    _dense("conv_fprop", 2.0 * N * H * W * Cin * Cout * R * S,
           lambda: lib.call("stc_conv_fprop_bnstats", x, wp, bias, y, N, H, W, Cin, Cout, R, S, dtype_code(x.dtype), config.engine, sums, ws,
                            ws.numel(), stream_ptr()))
    return y, sums


def conv_wgrad(x, dy, R: int, S: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    N, H, W, Cin = x.shape
    Cout = dy.shape[-1]
    wide = _widen(Cin, Cout, x.dtype, True, float(N) * H * W * Cin * Cout * R * S)
    if wide is not None:
        Cip, Cop = wide
        full = conv_wgrad(_resize_c(x, Cip), _resize_c(dy, Cop), R, S)          # (Cop, Cip, R, S)
        if out is None:
            out = torch.empty((Cout, Cin, R, S), dtype=torch.float32, device=x.device)
        out.copy_(full[:Cout, :Cin])
        return out
    ws = _wgrad_ws(R * S * Cin * Cout, x.device)
    _dense("conv_wgrad", 2.0 * N * H * W * Cin * Cout * R * S,
           lambda: lib.call("stc_conv_wgrad", x, dy, ws, N, H, W, Cin, Cout, R, S, dtype_code(x.dtype), config.engine, stream_ptr()))
    if out is None:
        out = torch.empty((Cout, Cin, R, S), dtype=torch.float32, device=x.device)
    lib.call("stc_unpack_conv_wgrad", ws, out, Cout, Cin, R, S, 0, stream_ptr())
    return out


def colsum(x2d_rows: int, C: int, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    if out is None:
        out = torch.empty(C, dtype=torch.float32, device=x.device)
    lib.call("stc_colsum", x, out, x2d_rows, C, 0, dtype_code(x.dtype), stream_ptr())
    return out


def add(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    out = torch.empty_like(a)
    lib.call("stc_add", a, b, out, a.numel(), dtype_code(a.dtype), stream_ptr())
    return out


def gemm(A, B, C, M, N, K, batch1, batch2, sA, sB, sC, alpha=1.0):
    d = GemmDesc(M, N, K, batch1, batch2, sA[0], sA[1], sA[2], sA[3], sB[0], sB[1], sB[2], sB[3], sC[0], sC[1], sC[2],
                 alpha, 0.0)
    _dense("gemm", 2.0 * M * N * K * batch1 * batch2,
           lambda: lib.call("stc_gemm", A, B, C, d, dtype_code(A.dtype), config.engine, stream_ptr()))


# ---------------------------------------------------------------------------------------------
# fan-out: explicit gradient summation with our add kernel (instead of autograd's implicit add)
# ---------------------------------------------------------------------------------------------
class _GradChain:
    """Running sum of the gradients that flow into ONE fanned-out tensor from consumers whose backward can add into it for free (a conv
    dgrad's residual input): each such consumer computes acc = its gradient + acc, returns None to autograd, and _Fanout.backward picks the sum
    up.  Replaces an add_n pass over n full-size tensors (KernelSelectAttention: 3 branch dgrads + the residual path)."""
    def __init__(self):
        self.acc = None


class _Fanout(Function):
    @staticmethod
    def forward(ctx, x, n):
        ctx.n = n
        ctx.set_materialize_grads(False)   # consumers that chain their gradient (see _GradChain) return None, not a tensor of zeros
        return tuple(x.view_as(x) for _ in range(n))

    @staticmethod
    def backward(ctx, *grads):
        gs = [_chk(g) for g in grads if g is not None]
        chain = getattr(ctx, "grad_chain", None)
        if chain is not None and chain.acc is not None:
            gs.append(chain.acc)
            chain.acc = None
        if not gs:
            return None, None
        acc = gs[0]
        i = 1
        while i < len(gs):  # up to three more addends per pass
            grp = gs[i:i + 3]
            out = torch.empty_like(acc)
            lib.call("stc_add_n", acc, grp[0], grp[1] if len(grp) > 1 else None, grp[2] if len(grp) > 2 else None, out,
                     acc.numel(), dtype_code(acc.dtype), stream_ptr())
            acc = out
            i += 3
        return acc, None


def fanout(x: torch.Tensor, n: int):
    if n == 1 or not (torch.is_grad_enabled() and x.requires_grad):
        return tuple(x for _ in range(n))
    return _Fanout.apply(x, n)


# ---------------------------------------------------------------------------------------------
# Conv + BatchNorm + activation  (DoubleConv halves, KSA branches, CoordAtt conv1+bn1+h_swish)
# ---------------------------------------------------------------------------------------------
class BNState:
    """Non-tensor side inputs of the BN part (buffers are updated in place by the kernels)."""

    def __init__(self, running_mean, running_var, num_batches_tracked, momentum, eps, training, sync=False, group=None):
        self.running_mean, self.running_var, self.nbt = running_mean, running_var, num_batches_tracked
        self.momentum, self.eps, self.training = momentum, eps, training
        self.sync, self.group = sync, group


def _bn_forward_stats(y, P, C, bn: BNState, sums=None):
    """sums: [sum y | sum y^2] (fp64, 2C) when the producing conv already reduced them in its epilogue (conv_fprop_bnstats)."""
    dev = y.device
    mean = torch.empty(C, dtype=torch.float32, device=dev)
    invstd = torch.empty(C, dtype=torch.float32, device=dev)
    count = float(P)
    if bn.training:
        if sums is None:
            sums = torch.empty(2 * C, dtype=torch.float64, device=dev)
            nb = lib.raw("stc_bn_ws_bytes")(P, C)
            ws = _workspace(dev, nb)
            lib.call("stc_bn_reduce", y, sums, P, C, ws, ws.numel(), dtype_code(y.dtype), stream_ptr())
        world = _sync_world(bn)
        if world > 1:
            # C1: one all-reduce of [sum, sumsq]; every rank holds the same per-GPU batch (count = P * world)
            _allreduce_stats(sums, bn)
            count = float(P) * world
        lib.call("stc_bn_finalize", sums, count, mean, invstd, bn.running_mean, bn.running_var, bn.nbt,
                 float(bn.momentum), float(bn.eps), C, stream_ptr())
    else:
        lib.call("stc_bn_eval_stats", bn.running_mean, bn.running_var, mean, invstd, float(bn.eps), C, stream_ptr())
    return mean, invstd, count


def _bn_backward(y, dout, mean, invstd, gamma, beta, P, C, act, bn, count, pg, pb, aff=None):
    """aff = (up_scale (N,C), up_shift (N,C), shift_scale, rows_per_image): the upstream gradient is up_scale*dout + up_shift*shift_scale
    per image (KSA branches, see ksa_fuse) and is never materialised."""
    training = bn.training
    dev = y.device
    sums = torch.empty(2 * C, dtype=torch.float64, device=dev)
    nb = lib.raw("stc_bn_ws_bytes")(P, C)
    ws = _workspace(dev, nb)
    code = dtype_code(y.dtype)
    if aff is not None:
        ua, ub, ubs, rows_img = aff
        n_img = P // rows_img
        lib.call("stc_bn_bwd_reduce_aff", y, dout, ua, ub, float(ubs), rows_img, n_img, mean, invstd, gamma, beta, sums, C, ws, ws.numel(),
                 code, stream_ptr())
    else:
        lib.call("stc_bn_bwd_reduce", y, dout, mean, invstd, gamma, beta, sums, P, C, act, ws, ws.numel(), code, stream_ptr())
    dgamma = _grad_buf(pg, (C,), dev)
    dbeta = _grad_buf(pb, (C,), dev)
    lib.call("stc_bn_param_grads", sums, dgamma, dbeta, C, stream_ptr())
    if training and _sync_world(bn) > 1:
        _allreduce_stats(sums, bn)  # C2
    dy = torch.empty_like(y)
    if aff is not None:
        lib.call("stc_bn_bwd_apply_aff", y, dout, ua, ub, float(ubs), rows_img, n_img, mean, invstd, gamma, beta, sums, float(count), dy, C,
                 code, stream_ptr())
    else:
        lib.call("stc_bn_bwd_apply", y, dout, mean, invstd, gamma, beta, sums, float(count), dy, P, C, act,
                 0 if training else 1, code, stream_ptr())
    return dy, dgamma, dbeta


def _check_channels(x: torch.Tensor, Cin: int, what: str):
    """The kernels take Cin from the activations: a layout mix-up (NHWC data read as NCHW, a wrong skip tensor) must raise, not read
    the packed weights out of bounds."""
    if x.dim() != 4 or x.shape[-1] != Cin:
        raise RuntimeError(f"{what}: activations are {tuple(x.shape)} (N, H, W, C) but the weight expects {Cin} input channels")


def _use_im2col(x, weight) -> bool:
    """Small-Cin convs (the image conv) go through im2col + a K=64 1x1 conv so they run on the tensor cores."""
    Cout, Cin, R, S = weight.shape
    return x.dtype == torch.bfloat16 and Cin < 8 and R * S * Cin <= 64 and Cout % 32 == 0 and config.engine != _lib.ENGINE_SIMT


def _im2col(x, R, S, Kpad=64):
    N, H, W, Cin = x.shape
    out = torch.empty((N, H, W, Kpad), dtype=x.dtype, device=x.device)
    lib.call("stc_im2col", x, out, N, H, W, Cin, R, S, Kpad, dtype_code(x.dtype), stream_ptr())
    return out


_EVAL_FOLD_CACHE: dict = {}     # (dtype, im2col, eps, (data_ptr, version) of every tensor involved) -> (packed folded weight, folded bias)


def _folded_eval_operands(weight, bias, gamma, beta, bn: BNState, dtype, im2col: bool):
    """(packed weight, bias) of the conv with the eval-mode BN folded in.  The folded + packed operand only changes when a parameter /
    running statistic does: cached on their version counters when config.cache_eval_weights."""
    Cout, Cin, R, S = weight.shape
    ts = (weight, gamma, beta, bn.running_mean, bn.running_var) + ((bias,) if bias is not None else ())
    key = (dtype, im2col, float(bn.eps)) + tuple((t.data_ptr(), t._version) for t in ts)
    hit = _EVAL_FOLD_CACHE.get(key) if config.cache_eval_weights else None
    if hit is None:
        wf = torch.empty_like(weight)
        bf = torch.empty(Cout, dtype=torch.float32, device=weight.device)
        lib.call("stc_bn_fold_conv", weight, bias, gamma, beta, bn.running_mean, bn.running_var, float(bn.eps), wf, bf, Cout, Cin * R * S,
                 stream_ptr())
        wp = pack_weight(wf, dtype, im2col_pad=64, cache=False) if im2col else pack_weight(wf, dtype, cache=False)
        hit = (wp, bf)
        if config.cache_eval_weights:
            if len(_EVAL_FOLD_CACHE) >= 1024:
                _EVAL_FOLD_CACHE.clear()
            _EVAL_FOLD_CACHE[key] = hit
    return hit


class _ConvBnAct(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, bn: BNState, act: int, pobjs):
        x = _chk(x)
        Cout, Cin, R, S = weight.shape
        _check_channels(x, Cin, "conv_bn_act")
        N, H, W, _ = x.shape
        P = N * H * W
        ctx.im2col = _use_im2col(x, weight)
        if not bn.training and not torch.is_grad_enabled() and config.fold_eval_bn:
            # inference: BN folded into the conv's weights / bias, activation in the conv epilogue - one pass instead of three
            # the folded + packed operand only changes when a parameter / running statistic does: cached on their version counters
            hit = _folded_eval_operands(weight, bias, gamma, beta, bn, x.dtype, ctx.im2col)
            wp, bf = hit
            if ctx.im2col:
                return conv_fprop(_im2col(x, R, S), wp, bf, None, Cout, 1, 1, act)
            return conv_fprop(x, wp, bf, None, Cout, R, S, act)
        if ctx.im2col:
            x = _im2col(x, R, S)          # (N,H,W,64): saved instead of the 3-channel image for the wgrad GEMM
            wp = pack_weight(weight, x.dtype, im2col_pad=64)
            y = conv_fprop(x, wp, bias, None, Cout, 1, 1)
            sums = None
        elif bn.training and lib.raw("stc_conv_bnstats_fused_ok")(W, Cin, Cout, R, S, dtype_code(x.dtype), config.engine):
            # conv + batch statistics in one pass: the halo kernel's epilogue reduces sum / sum of squares of what it stores
            wp = pack_weight(weight, x.dtype)
            y, sums = conv_fprop_bnstats(x, wp, bias, Cout, R, S)
        else:
            wp = pack_weight(weight, x.dtype)
            y = conv_fprop(x, wp, bias, None, Cout, R, S)
            sums = None
        mean, invstd, count = _bn_forward_stats(y, P, Cout, bn, sums)
        a = torch.empty_like(y)
        lib.call("stc_bn_apply", y, mean, invstd, gamma, beta, a, P, Cout, act, dtype_code(y.dtype), stream_ptr())
        ctx.save_for_backward(x, y, weight, gamma, beta, mean, invstd)
        ctx.meta = (act, bn, count, R, S, pobjs, bias is not None)
        return a

    @staticmethod
    def backward(ctx, da):
        x, y, weight, gamma, beta, mean, invstd = ctx.saved_tensors
        act, bn, count, R, S, pobjs, has_bias = ctx.meta
        pw, pbias, pg, pb = pobjs
        da = _chk(da)
        N, H, W, Cout = y.shape
        P = N * H * W
        aff = None
        slot = getattr(ctx, "ksa_slot", None)      # set by ksa_fuse: `da` is KSA's raw dout, this branch's gradient is affine in it
        if slot is not None and slot[0].ready:
            box, k = slot
            aff = (box.scale[k], box.shift, box.shift_scale, box.rows)
        dy, dgamma, dbeta = _bn_backward(y, da, mean, invstd, gamma, beta, P, Cout, act, bn, count, pg, pb, aff)
        dbias = None
        if has_bias:
            if bn.training:
                # sum_p dy == gamma*invstd*(sum g - n*mean(g) - sum(xhat)*mean(g*xhat)) == 0 exactly: a conv bias feeding a
                # train-mode BN has no gradient (the reference's autograd returns fp32 rounding noise here)
                dbias = _grad_buf(pbias, (Cout,), dy.device, zero=True)
            else:
                dbias = colsum(P, Cout, dy, _grad_buf(pbias, (Cout,), dy.device))
        if ctx.im2col:
            if ctx.needs_input_grad[0]:
                raise RuntimeError("im2col conv path does not provide an input gradient (it is only used for the image conv)")
            ws = _wgrad_ws(64 * Cout, dy.device)
            _dense("conv_wgrad", 2.0 * P * 64 * Cout,
                   lambda: lib.call("stc_conv_wgrad", x, dy, ws, N, H, W, 64, Cout, 1, 1, dtype_code(x.dtype), config.engine, stream_ptr()))
            dw = _grad_buf(pw, weight.shape, dy.device)
            lib.call("stc_unpack_im2col_wgrad", ws, dw, Cout, weight.shape[1], R, S, stream_ptr())
            return None, dw, dbias, dgamma, dbeta, None, None, None
        xw = x
        if R == 1 and S == 1 and bn.training and x.dtype == torch.float32 and _sync_world(bn) == 1:
            # exact-parity path: the rows of dy sum to zero (train-mode BN), so dW = dy^T x = dy^T (x - 1 c^T) for ANY c; with
            # c = the column mean the contraction no longer multiplies the rounding residue of sum(dy) (and every dy rounding
            # error) by the common offset of x — CoordAtt's conv1 sees descriptors ~100x larger than their fluctuation
            xw = _center_tokens(x.view(1, N * H * W, x.shape[-1])).view_as(x)
        dw = conv_wgrad(xw, dy, R, S, _grad_buf(pw, weight.shape, dy.device))
        dx = None
        if ctx.needs_input_grad[0]:
            wpt = pack_weight(weight, dy.dtype, transpose_flip=True)
            chain = getattr(ctx, "grad_chain", None)   # set by ksa_fuse: the input's other gradients are added in the dgrad epilogue
            dx = conv_fprop(dy, wpt, None, chain.acc if chain is not None else None, weight.shape[1], R, S)
            if chain is not None:
                chain.acc, dx = dx, None
        return dx, dw, dbias, dgamma, dbeta, None, None, None


def _bn_state(bn, training: bool) -> BNState:
    sync = isinstance(bn, torch.nn.SyncBatchNorm)
    return BNState(bn.running_mean, bn.running_var, bn.num_batches_tracked, 0.1 if bn.momentum is None else bn.momentum, bn.eps,
                   training or not bn.track_running_stats, sync=sync, group=getattr(bn, "process_group", None) if sync else None)


# ---------------------------------------------------------------------------------------------
# Conv + BN + activation over a VIRTUAL channel concat: the conv reads its sources directly, the concatenated
# tensor of Up.forward (unet_head.py:54-55) / UpConvBlock.forward (up_conv_block.py:99) / the UNet++ decoder blocks is never written
# ---------------------------------------------------------------------------------------------
def _pad5(vals, fill):
    vals = list(vals)
    return vals + [fill] * (5 - len(vals))


def cat_ok(cins, Cout: int, dtype, needs_dgrad: bool = True) -> bool:
    """True when the tcgen05 kernels take this list of sources as a virtual concat (bf16, every part a multiple of 64 channels)."""
    if dtype != torch.bfloat16 or not 2 <= len(cins) <= 5:
        return False
    return bool(lib.raw("stc_conv_cat_ok")(*_pad5(cins, 0), int(Cout), dtype_code(dtype), config.engine))


def conv_fprop_cat(xs, wp, bias, Cout: int, R: int, S: int, act: int = 0) -> torch.Tensor:
    N, H, W, _ = xs[0].shape
    cins = [x.shape[-1] for x in xs]
    y = torch.empty((N, H, W, Cout), dtype=xs[0].dtype, device=xs[0].device)
    _dense("conv_fprop", 2.0 * N * H * W * sum(cins) * Cout * R * S,
           lambda: lib.call("stc_conv_fprop_cat", *_pad5(xs, None), *_pad5(cins, 0), wp, bias, y, N, H, W, Cout, R, S, act, dtype_code(y.dtype),
                            config.engine, stream_ptr()))
    return y


def conv_wgrad_cat(xs, dy, R: int, S: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    N, H, W, _ = xs[0].shape
    cins = [x.shape[-1] for x in xs]
    Cin, Cout = sum(cins), dy.shape[-1]
    ws = _wgrad_ws(R * S * Cin * Cout, dy.device)
    _dense("conv_wgrad", 2.0 * N * H * W * Cin * Cout * R * S,
           lambda: lib.call("stc_conv_wgrad_cat", *_pad5(xs, None), *_pad5(cins, 0), dy, ws, N, H, W, Cout, R, S, dtype_code(dy.dtype),
                            config.engine, stream_ptr()))
    if out is None:
        out = torch.empty((Cout, Cin, R, S), dtype=torch.float32, device=dy.device)
    lib.call("stc_unpack_conv_wgrad", ws, out, Cout, Cin, R, S, 0, stream_ptr())
    return out


def conv_dgrad_split(dy, wpt, cins, R: int, S: int):
    """Input gradients of a conv over a virtual concat: one tensor per source, written by ONE dgrad launch."""
    N, H, W, Cdy = dy.shape
    dxs = [torch.empty((N, H, W, c), dtype=dy.dtype, device=dy.device) for c in cins]
    _dense("conv_fprop", 2.0 * N * H * W * sum(cins) * Cdy * R * S,
           lambda: lib.call("stc_conv_dgrad_split", dy, wpt, *_pad5(dxs, None), *_pad5(cins, 0), N, H, W, Cdy, R, S, dtype_code(dy.dtype),
                            config.engine, stream_ptr()))
    return dxs


class _ConvBnActCat(Function):
    @staticmethod
    def forward(ctx, weight, bias, gamma, beta, bn: BNState, act: int, pobjs, *xs):
        xs = [_chk(x) for x in xs]
        Cout, Cin, R, S = weight.shape
        cins = [x.shape[-1] for x in xs]
        if sum(cins) != Cin or any(x.shape[:3] != xs[0].shape[:3] for x in xs):
            raise RuntimeError(f"conv_bn_act_cat: sources {[tuple(x.shape) for x in xs]} do not concatenate to {Cin} input channels")
        N, H, W, _ = xs[0].shape
        P = N * H * W
        if not bn.training and not torch.is_grad_enabled() and config.fold_eval_bn:
            wp, bf = _folded_eval_operands(weight, bias, gamma, beta, bn, xs[0].dtype, False)
            return conv_fprop_cat(xs, wp, bf, Cout, R, S, act)
        wp = pack_weight(weight, xs[0].dtype)
        y = conv_fprop_cat(xs, wp, bias, Cout, R, S)
        mean, invstd, count = _bn_forward_stats(y, P, Cout, bn, None)
        a = torch.empty_like(y)
        lib.call("stc_bn_apply", y, mean, invstd, gamma, beta, a, P, Cout, act, dtype_code(y.dtype), stream_ptr())
        ctx.save_for_backward(y, weight, gamma, beta, mean, invstd, *xs)
        ctx.meta = (act, bn, count, R, S, pobjs, bias is not None, cins)
        return a

    @staticmethod
    def backward(ctx, da):
        y, weight, gamma, beta, mean, invstd, *xs = ctx.saved_tensors
        act, bn, count, R, S, pobjs, has_bias, cins = ctx.meta
        pw, pbias, pg, pb = pobjs
        da = _chk(da)
        N, H, W, Cout = y.shape
        P = N * H * W
        dy, dgamma, dbeta = _bn_backward(y, da, mean, invstd, gamma, beta, P, Cout, act, bn, count, pg, pb, None)
        dbias = None
        if has_bias:
            dbias = _grad_buf(pbias, (Cout,), dy.device, zero=True) if bn.training else colsum(P, Cout, dy, _grad_buf(pbias, (Cout,), dy.device))
        dw = conv_wgrad_cat(xs, dy, R, S, _grad_buf(pw, weight.shape, dy.device))
        dxs = [None] * len(xs)
        if any(ctx.needs_input_grad[7:]):
            wpt = pack_weight(weight, dy.dtype, transpose_flip=True)
            got = conv_dgrad_split(dy, wpt, cins, R, S)
            dxs = [g if need else None for g, need in zip(got, ctx.needs_input_grad[7:])]
        return (dw, dbias, dgamma, dbeta, None, None, None, *dxs)


def conv_bn_act_cat(xs, conv: torch.nn.Conv2d, bn: torch.nn.modules.batchnorm._BatchNorm, act: int, training: bool, materialise=None):
    """conv_bn_act(cat(xs, channels)) without writing the concatenation when the tcgen05 kernels take the sources directly; otherwise
    `materialise(xs)` (default: ops.cat_channels_n) builds the concatenated tensor for the ordinary path."""
    xs = list(xs)
    Cout = conv.weight.shape[0]
    if len(xs) >= 2 and cat_ok([x.shape[-1] for x in xs], Cout, xs[0].dtype) and all(x.shape[:3] == xs[0].shape[:3] for x in xs):
        return _ConvBnActCat.apply(conv.weight, conv.bias, bn.weight, bn.bias, _bn_state(bn, training), act,
                                   (conv.weight, conv.bias, bn.weight, bn.bias), *xs)
    x = xs[0] if len(xs) == 1 else (materialise(xs) if materialise is not None else cat_channels_n(xs))
    return conv_bn_act(x, conv, bn, act, training)


def conv_bn_act(x, conv: torch.nn.Conv2d, bn: torch.nn.modules.batchnorm._BatchNorm, act: int, training: bool):
    sync = isinstance(bn, torch.nn.SyncBatchNorm)
    state = BNState(bn.running_mean, bn.running_var, bn.num_batches_tracked,
                    0.1 if bn.momentum is None else bn.momentum, bn.eps, training or not bn.track_running_stats,
                    sync=sync, group=getattr(bn, "process_group", None) if sync else None)
    return _ConvBnAct.apply(x, conv.weight, conv.bias, bn.weight, bn.bias, state, act,
                            (conv.weight, conv.bias, bn.weight, bn.bias))


# ---------------------------------------------------------------------------------------------
# Conv / Linear without norm: y = act(conv(x) + bias + residual)
# ---------------------------------------------------------------------------------------------
class _Conv(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, residual, act: int, pobjs):
        x = _chk(x)
        if residual is not None:
            residual = _chk(residual)
        Cout, Cin, R, S = weight.shape
        _check_channels(x, Cin, "conv2d / linear_tokens")
        ctx.im2col = _use_im2col(x, weight)
        if ctx.im2col:   # small-Cin conv (VGG16's first conv in UNet++): K=64 1x1 conv on the tensor cores
            x = _im2col(x, R, S)
            y = conv_fprop(x, pack_weight(weight, x.dtype, im2col_pad=64), bias, residual, Cout, 1, 1, act)
        else:
            wp = pack_weight(weight, x.dtype, cache=len(pobjs) < 3 or not pobjs[2])
            y = conv_fprop(x, wp, bias, residual, Cout, R, S, act)
        ctx.save_for_backward(x, weight, y if act != 0 else None)
        ctx.meta = (act, R, S, pobjs, bias is not None, residual is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        act, R, S, pobjs, has_bias, has_res = ctx.meta
        pw, pbias = pobjs[0], pobjs[1]
        dy = _chk(dy)
        if act != 0:
            dz = torch.empty_like(dy)
            lib.call("stc_act_bwd", y, dy, dz, dy.numel(), act, dtype_code(dy.dtype), stream_ptr())
        else:
            dz = dy
        Cout = weight.shape[0]
        P = dz.numel() // Cout
        dbias = colsum(P, Cout, dz, _grad_buf(pbias, (Cout,), dz.device)) if has_bias else None
        if ctx.im2col:
            if ctx.needs_input_grad[0]:
                raise RuntimeError("im2col conv path does not provide an input gradient (image convs only)")
            N_, H_, W_, _ = x.shape
            ws = _wgrad_ws(64 * Cout, dz.device)
            _dense("conv_wgrad", 2.0 * P * 64 * Cout,
                   lambda: lib.call("stc_conv_wgrad", x, dz, ws, N_, H_, W_, 64, Cout, 1, 1, dtype_code(x.dtype), config.engine, stream_ptr()))
            dw = _grad_buf(pw, weight.shape, dz.device)
            lib.call("stc_unpack_im2col_wgrad", ws, dw, Cout, weight.shape[1], R, S, stream_ptr())
            return None, dw, dbias, (dz if (has_res and ctx.needs_input_grad[3]) else None), None, None
        dw = conv_wgrad(x, dz, R, S, _grad_buf(pw, weight.shape, dz.device)) if ctx.needs_input_grad[1] else None
        dx = None
        if ctx.needs_input_grad[0]:
            wpt = pack_weight(weight, dz.dtype, transpose_flip=True, cache=len(pobjs) < 3 or not pobjs[2])
            dx = conv_fprop(dz, wpt, None, None, weight.shape[1], R, S)
        dres = dz if (has_res and ctx.needs_input_grad[3]) else None
        return dx, dw, dbias, dres, None, None


def conv2d(x, weight, bias=None, residual=None, act: int = 0, pobjs=None):
    """x NHWC; weight OIHW (odd kernel, padding k//2)."""
    return _Conv.apply(x, weight, bias, residual, act, pobjs or (weight, bias))


def linear_tokens(x, weight, bias=None, residual=None, act: int = 0, pobjs=None):
    """x (..., in) -> (..., out) with weight (out, in): a 1x1 conv over a (1,1,rows,in) image."""
    shp = x.shape
    rows = x.numel() // shp[-1]
    y = _Conv.apply(x.reshape(1, 1, rows, shp[-1]), weight.view(weight.shape[0], weight.shape[1], 1, 1), bias,
                    None if residual is None else residual.reshape(1, 1, rows, weight.shape[0]), act,
                    pobjs or (weight, bias))
    return y.view(*shp[:-1], weight.shape[0])


# ---------------------------------------------------------------------------------------------
# MaxPool2d(2)
# ---------------------------------------------------------------------------------------------
class _MaxPool2(Function):
    @staticmethod
    def forward(ctx, x):
        x = _chk(x)
        N, H, W, C = x.shape
        y = torch.empty((N, H // 2, W // 2, C), dtype=x.dtype, device=x.device)
        lib.call("stc_maxpool2_fwd", x, y, N, H, W, C, dtype_code(x.dtype), stream_ptr())
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = _chk(dy)
        N, H, W, C = x.shape
        dx = torch.empty_like(x)
        lib.call("stc_maxpool2_bwd", x, dy, dx, N, H, W, C, dtype_code(x.dtype), stream_ptr())
        return dx


def maxpool2(x):
    return _MaxPool2.apply(x)


# ---------------------------------------------------------------------------------------------
# bilinear x2 upsample + pad + concat
# ---------------------------------------------------------------------------------------------
class _UpCat(Function):
    @staticmethod
    def forward(ctx, skip, low, align_corners: bool):
        skip, low = _chk(skip), _chk(low)
        N, H, W, Cs = skip.shape
        _, h, w, Cu = low.shape
        out = torch.empty((N, H, W, Cs + Cu), dtype=skip.dtype, device=skip.device)
        lib.call("stc_upcat_fwd", skip, low, out, N, H, W, Cs, h, w, Cu, int(align_corners), dtype_code(skip.dtype), stream_ptr())
        ctx.meta = (N, H, W, Cs, h, w, Cu, int(align_corners))
        return out

    @staticmethod
    def backward(ctx, dout):
        N, H, W, Cs, h, w, Cu, ac = ctx.meta
        dout = _chk(dout)
        dskip = torch.empty((N, H, W, Cs), dtype=dout.dtype, device=dout.device) if ctx.needs_input_grad[0] else None
        dlow = torch.empty((N, h, w, Cu), dtype=dout.dtype, device=dout.device) if ctx.needs_input_grad[1] else None
        lib.call("stc_upcat_bwd", dout, dskip, dlow, N, H, W, Cs, h, w, Cu, ac, dtype_code(dout.dtype), stream_ptr())
        return dskip, dlow, None


def upcat(skip, low, align_corners=True):
    return _UpCat.apply(skip, low, align_corners)


# ---------------------------------------------------------------------------------------------
# Up.forward with CoordAtt, fused: cat = [skip, up(low)] is never materialised (csrc/upcat_fused.cu)
# ---------------------------------------------------------------------------------------------
class _UpCatShared:
    """Hand-over between the two nodes below: _UpCatApply.backward leaves dout here, _UpCatPool.backward (which autograd can only
    run afterwards: its dy depends on da) folds it with dy into ONE pass that writes dskip and dlow."""

    def __init__(self):
        self.dout = None


class _UpCatPool(Function):
    @staticmethod
    def forward(ctx, skip, low, align_corners: bool, shared: _UpCatShared):
        skip, low = _chk(skip), _chk(low)
        N, H, W, Cs = skip.shape
        _, h, w, Cu = low.shape
        y = torch.empty((N, H + W, Cs + Cu), dtype=skip.dtype, device=skip.device)
        ws = _workspace(skip.device, lib.raw("stc_upcat_pool_ws_bytes")(N, h, w, Cu))
        lib.call("stc_upcat_pool", skip, low, y, N, H, W, Cs, h, w, Cu, int(align_corners), ws, ws.numel(), dtype_code(skip.dtype), stream_ptr())
        ctx.meta = (N, H, W, Cs, h, w, Cu, int(align_corners), shared, skip.dtype, skip.device)
        ctx.set_materialize_grads(False)
        return y, skip.view_as(skip), low.view_as(low)

    @staticmethod
    def backward(ctx, dy, gs, gl):
        N, H, W, Cs, h, w, Cu, ac, shared, dt, dev = ctx.meta
        assert gs is None and gl is None, "the skip/low aliases of upcat_coordatt must only feed its apply node"
        dout, shared.dout = shared.dout, None
        if dout is None and dy is None:
            return None, None, None, None
        if dout is None:
            dout = torch.zeros((N, H, W, Cs + Cu), dtype=dt, device=dev)
        dskip = torch.empty((N, H, W, Cs), dtype=dt, device=dev) if ctx.needs_input_grad[0] else None
        dlow = torch.empty((N, h, w, Cu), dtype=dt, device=dev) if ctx.needs_input_grad[1] else None
        lib.call("stc_upcat_apply_bwd", dout, None if dy is None else _chk(dy), dskip, dlow, N, H, W, Cs, h, w, Cu, ac, dtype_code(dt),
                 stream_ptr())
        return dskip, dlow, None, None


class _UpCatApply(Function):
    @staticmethod
    def forward(ctx, skip, low, a, align_corners: bool, shared: _UpCatShared):
        N, H, W, Cs = skip.shape
        _, h, w, Cu = low.shape
        a = _chk(a)
        out = torch.empty((N, H, W, Cs + Cu), dtype=skip.dtype, device=skip.device)
        lib.call("stc_upcat_apply_fwd", skip, low, a, out, N, H, W, Cs, h, w, Cu, int(align_corners), dtype_code(skip.dtype), stream_ptr())
        ctx.save_for_backward(a)
        ctx.meta = (N, H, W, Cs + Cu, shared)
        return out

    @staticmethod
    def backward(ctx, dout):
        (a,) = ctx.saved_tensors
        N, H, W, C, shared = ctx.meta
        dout = _chk(dout)
        shared.dout = dout
        da = None
        if ctx.needs_input_grad[2]:
            da = torch.empty_like(a)
            lib.call("stc_coordatt_apply_bwd", dout, a, da, N, H, W, C, dtype_code(dout.dtype), stream_ptr())
        return None, None, da, None, None


def upcat_coordatt(skip, low, align_corners, attention):
    """out = cat + a_h*a_w with cat = [skip, pad(bilinear_x2(low))] and a = attention(row/col means of cat)  (Up.forward with se=True).
    Falls back to upcat + coordatt_pool/apply for channel counts the fused kernels do not take."""
    N, H, W, Cs = skip.shape
    _, h, w, Cu = low.shape
    if not lib.raw("stc_upcat_fused_ok")(N, H, W, Cs, h, w, Cu):
        y, xb = coordatt_pool(upcat(skip, low, align_corners))
        return coordatt_apply(xb, attention(y, N, H, W))
    shared = _UpCatShared()
    y, s_alias, l_alias = _UpCatPool.apply(skip, low, align_corners, shared)
    return _UpCatApply.apply(s_alias, l_alias, attention(y, N, H, W), align_corners, shared)


# ---------------------------------------------------------------------------------------------
# CoordAtt: pooled descriptors and the additive attention map
# ---------------------------------------------------------------------------------------------
class _CoordAttPool(Function):
    """x -> (y, x_alias): y = row/col means (N,H+W,C); x_alias is x itself, to be consumed by coordatt_apply.  Having both
    consumers of x behind ONE node lets backward form dx = d(x_alias) + dy_h/W + dy_w/H in a single pass."""

    @staticmethod
    def forward(ctx, x):
        x = _chk(x)
        N, H, W, C = x.shape
        y = torch.empty((N, H + W, C), dtype=x.dtype, device=x.device)
        lib.call("stc_rowcol_mean", x, y, N, H, W, C, dtype_code(x.dtype), stream_ptr())
        ctx.shape = (N, H, W, C)
        return y, x.view_as(x)

    @staticmethod
    def backward(ctx, dy, dxa):
        N, H, W, C = ctx.shape
        if dy is None:
            return dxa
        dy = _chk(dy)
        if dxa is None:
            dxa = torch.zeros((N, H, W, C), dtype=dy.dtype, device=dy.device)
        dxa = _chk(dxa)
        dx = torch.empty_like(dxa)
        lib.call("stc_coordatt_dx", dxa, dy, dx, N, H, W, C, dtype_code(dy.dtype), stream_ptr())
        return dx


def coordatt_pool(x):
    return _CoordAttPool.apply(x)


class _RowColMean(Function):
    """(N,H,W,C) -> (N,H+W,C): rows 0..H-1 = mean over W, rows H.. = mean over H.  Its backward is folded
    into _CoordAttApply (which owns the only other use of x), so this node returns no grad for x."""

    @staticmethod
    def forward(ctx, x):
        x = _chk(x)
        N, H, W, C = x.shape
        y = torch.empty((N, H + W, C), dtype=x.dtype, device=x.device)
        lib.call("stc_rowcol_mean", x, y, N, H, W, C, dtype_code(x.dtype), stream_ptr())
        ctx.shape = (N, H, W, C)
        return y

    @staticmethod
    def backward(ctx, dy):
        N, H, W, C = ctx.shape
        dy = _chk(dy)
        dx = torch.zeros((N, H, W, C), dtype=dy.dtype, device=dy.device)
        lib.call("stc_coordatt_dx", dx, dy, dx, N, H, W, C, dtype_code(dy.dtype), stream_ptr())
        return dx


class _CoordAttApply(Function):
    """out = x + a_h * a_w with a = (N,H+W,C)."""

    @staticmethod
    def forward(ctx, x, a):
        x, a = _chk(x), _chk(a)
        N, H, W, C = x.shape
        out = torch.empty_like(x)
        lib.call("stc_coordatt_apply", x, a, out, N, H, W, C, dtype_code(x.dtype), stream_ptr())
        ctx.save_for_backward(a)
        ctx.shape = (N, H, W, C)
        return out

    @staticmethod
    def backward(ctx, dout):
        (a,) = ctx.saved_tensors
        N, H, W, C = ctx.shape
        dout = _chk(dout)
        da = torch.empty_like(a)
        lib.call("stc_coordatt_apply_bwd", dout, a, da, N, H, W, C, dtype_code(dout.dtype), stream_ptr())
        return dout, da


def rowcol_mean(x):
    return _RowColMean.apply(x)


def coordatt_apply(x, a):
    return _CoordAttApply.apply(x, a)


# ---------------------------------------------------------------------------------------------
# KernelSelectAttention fuse: out = x + sum_k softmax_k(fcs_k(fc(GAP(f0+f1+f2)))) * f_k
# ---------------------------------------------------------------------------------------------
class _KSALazy:
    """Hand-over between _KSAFuse.backward and the three branch _ConvBnAct.backward nodes: df_k = scale[k] * dout + shift * shift_scale."""
    def __init__(self):
        self.ready, self.scale, self.shift, self.shift_scale, self.rows = False, None, None, 0.0, 0


class _KSAFuse(Function):
    @staticmethod
    def forward(ctx, x, f0, f1, f2, fc_w, fc_b, w0, b0, w1, b1, w2, b2, pobjs, lazy=None, chain=None):
        ctx.lazy = lazy
        ctx.chain = chain
        x, f0, f1, f2 = _chk(x), _chk(f0), _chk(f1), _chk(f2)
        N, H, W, C = x.shape
        HW = H * W
        dev = x.device
        d = fc_w.shape[0]
        code = dtype_code(x.dtype)
        S = torch.empty((N, C), dtype=torch.float32, device=dev)
        lib.call("stc_ksa_pool", f0, f1, f2, S, N, HW, C, code, stream_ptr())
        Z = torch.empty((N, d), dtype=torch.float32, device=dev)
        lib.call("stc_linear_f32_fwd", S, fc_w, fc_b, Z, N, C, d, stream_ptr())
        a = torch.empty((3, N, C), dtype=torch.float32, device=dev)
        for k, (wk, bk) in enumerate(((w0, b0), (w1, b1), (w2, b2))):
            lib.call("stc_linear_f32_fwd", Z, wk, bk, a[k], N, d, C, stream_ptr())
        wts = torch.empty_like(a)
        lib.call("stc_softmax3_fwd", a, wts, N * C, stream_ptr())
        out = torch.empty_like(x)
        lib.call("stc_ksa_combine", x, f0, f1, f2, wts, out, N, HW, C, code, stream_ptr())
        ctx.save_for_backward(f0, f1, f2, wts, S, Z, fc_w, w0, w1, w2)
        ctx.pobjs = pobjs
        return out

    @staticmethod
    def backward(ctx, dout):
        f0, f1, f2, wts, S, Z, fc_w, w0, w1, w2 = ctx.saved_tensors
        pobjs = ctx.pobjs
        dout = _chk(dout)
        N, H, W, C = dout.shape
        HW = H * W
        dev = dout.device
        d = fc_w.shape[0]
        code = dtype_code(dout.dtype)
        dw = torch.empty((3, N, C), dtype=torch.float32, device=dev)
        lib.call("stc_ksa_dw", dout, f0, f1, f2, dw, N, HW, C, code, stream_ptr())
        da = torch.empty_like(dw)
        lib.call("stc_softmax3_bwd", wts, dw, da, N * C, stream_ptr())
        dZ = torch.zeros((N, d), dtype=torch.float32, device=dev)
        dZk = torch.empty_like(dZ)
        grads = []
        for k, wk in enumerate((w0, w1, w2)):
            gW = _grad_buf(pobjs[2 + 2 * k], wk.shape, dev, zero=True)
            gb = _grad_buf(pobjs[3 + 2 * k], (C,), dev, zero=True)
            lib.call("stc_linear_f32_bwd", Z, wk, da[k], dZk, gW, gb, N, d, C, stream_ptr())
            lib.call("stc_axpy_f32", dZk, dZ, 1.0, dZ.numel(), stream_ptr())
            grads += [gW, gb]
        dS = torch.empty((N, C), dtype=torch.float32, device=dev)
        gfcW = _grad_buf(pobjs[0], fc_w.shape, dev, zero=True)
        gfcb = _grad_buf(pobjs[1], (d,), dev, zero=True)
        lib.call("stc_linear_f32_bwd", S, fc_w, dZ, dS, gfcW, gfcb, N, C, d, stream_ptr())
        if ctx.lazy is not None:
            # the branch BN backward kernels read dout themselves (stc_bn_bwd_*_aff): no df tensors (4 x |x| of traffic less per level)
            lz = ctx.lazy
            lz.scale, lz.shift, lz.shift_scale, lz.rows, lz.ready = wts, dS, 1.0 / HW, HW, True
            dx = dout
            if ctx.chain is not None:   # the residual path's gradient starts the chain the three branch dgrads add to (read, never written)
                ctx.chain.acc, dx = dout, None
            return (dx, dout, dout, dout, gfcW, gfcb, *grads, None, None, None)
        df0, df1, df2 = torch.empty_like(f0), torch.empty_like(f1), torch.empty_like(f2)
        lib.call("stc_ksa_df", dout, wts, dS, df0, df1, df2, N, HW, C, code, stream_ptr())
        return (dout, df0, df1, df2, gfcW, gfcb, *grads, None, None, None)


def ksa_fuse(x, f0, f1, f2, fc, fcs):
    ps = (fc.weight, fc.bias, fcs[0].weight, fcs[0].bias, fcs[1].weight, fcs[1].bias, fcs[2].weight, fcs[2].bias)
    lazy = None
    fns = [f.grad_fn for f in (f0, f1, f2)]
    C = x.shape[-1]
    if (config.ksa_lazy_df and all(fn is not None and type(fn).__name__ == "_ConvBnActBackward" and fn.meta[0] == _lib.ACT_RELU
                                   and fn.meta[1].training for fn in fns)
            and len({id(fn) for fn in fns}) == 3 and lib.raw("stc_bn_bwd_aff_ok")(C)):
        lazy = _KSALazy()
        for k, fn in enumerate(fns):
            fn.ksa_slot = (lazy, k)
    chain = None
    fan = x.grad_fn
    if (lazy is not None and config.chain_fanout_grads and fan is not None and type(fan).__name__ == "_FanoutBackward"
            and all(fn.next_functions[0][0] is fan for fn in fns)):
        # x and the three branch inputs are the four outputs of one fan-out: their gradients are summed inside the branch dgrads
        chain = _GradChain()
        fan.grad_chain = chain
        for fn in fns:
            fn.grad_chain = chain
    return _KSAFuse.apply(x, f0, f1, f2, *ps, ps, lazy, chain)


# ---------------------------------------------------------------------------------------------
# Multi-head attention core: softmax(Q K^T / sqrt(hd)) V on (N, L, E) token tensors
# ---------------------------------------------------------------------------------------------
def _softmax_gemm(a, b, P, dS, L, hd, N, heads, sA, sB, sC, scale) -> bool:
    """Row softmax inside the score product (stc_gemm_softmax / stc_gemm_softmax_bwd: two sweeps over a row block's N tiles, no L x L score
    or dP tensor).  dS is None: P = softmax(scale * a b^T) is written into P.  Else: dS = scale * P * (a b^T - sum(P dP) / sum(P)).
    False when the tcgen05 engine does not take the shape (the caller then runs the product and the softmax pass separately)."""
    if not config.fuse_attention_softmax or a.dtype != torch.bfloat16:
        return False
    d = GemmDesc(L, L, hd, N, heads, sA[0], sA[1], sA[2], sA[3], sB[0], sB[1], sB[2], sB[3], sC[0], sC[1], sC[2], 1.0, 0.0)
    out = P if dS is None else dS
    if not lib.raw("stc_gemm_softmax_ok")(ctypes.byref(d), ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(P.data_ptr()), dtype_code(a.dtype), config.engine):
        return False
    flops = 2.0 * L * L * hd * N * heads   # algorithmic: ONE product (the second sweep's recomputation is this kernel's choice, not work)
    if dS is None:
        _dense("gemm", flops, lambda: lib.call("stc_gemm_softmax", a, b, P, d, float(scale), dtype_code(a.dtype), config.engine, stream_ptr()))
    else:
        _dense("gemm", flops, lambda: lib.call("stc_gemm_softmax_bwd", a, b, P, dS, d, float(scale), dtype_code(a.dtype), config.engine, stream_ptr()))
    return True


def _dsoftmax_gemm(do, v, P, o, dS, L, hd, N, heads, sA, sB, sC, scale) -> bool:
    """dS = scale * P * (dO V^T - rowsum(dO * O)) out of the dP product's epilogue (stc_gemm_dsoftmax); False when the tcgen05 engine
    does not take it (the caller then runs the product and the softmax-backward pass separately)."""
    d = GemmDesc(L, L, hd, N, heads, sA[0], sA[1], sA[2], sA[3], sB[0], sB[1], sB[2], sB[3], sC[0], sC[1], sC[2], float(scale), 0.0)
    if not lib.raw("stc_gemm_dsoftmax_ok")(ctypes.byref(d), dtype_code(do.dtype), config.engine):
        return False
    D = torch.empty((N, heads, L), dtype=torch.float32, device=do.device)
    lib.call("stc_rowdot_heads", do, o, D, N, L, heads, hd, dtype_code(do.dtype), stream_ptr())
    _dense("gemm", 2.0 * L * L * hd * N * heads,
           lambda: lib.call("stc_gemm_dsoftmax", do, v, P, D, dS, d, dtype_code(do.dtype), config.engine, stream_ptr()))
    return True


def _center_tokens(x: torch.Tensor) -> torch.Tensor:
    """x (N, L, E) -> x - mean over the L tokens (stc_center_tokens)."""
    N, L, E = x.shape
    out = torch.empty_like(x)
    ws = torch.empty(N * E, dtype=torch.float32, device=x.device)
    lib.call("stc_center_tokens", x, out, ws, N, L, E, dtype_code(x.dtype), stream_ptr())
    return out


class _Attention(Function):
    @staticmethod
    def forward(ctx, q, k, v, heads: int):
        q, k, v = _chk(q), _chk(k), _chk(v)
        N, L, E = q.shape
        hd = E // heads
        dev = q.device
        P = torch.empty((N, heads, L, L), dtype=q.dtype, device=dev)
        tok = (L * E, hd)  # batch strides of a (N, L, E) tensor split into heads
        if q.dtype == torch.float32:   # exact-parity path: centred keys (same softmax, no common component in the fp32 scores)
            k = _center_tokens(k)
        scale = 1.0 / math.sqrt(hd)
        if not _softmax_gemm(q, k, P, None, L, hd, N, heads, (*tok, E, 1), (*tok, 1, E), (heads * L * L, L * L, L), scale):
            gemm(q, k, P, L, L, hd, N, heads, (*tok, E, 1), (*tok, 1, E), (heads * L * L, L * L, L))
            lib.call("stc_softmax_rows_fwd", P, P, N * heads * L, L, scale, dtype_code(q.dtype), stream_ptr())
        o = torch.empty_like(q)
        gemm(P, v, o, L, hd, L, N, heads, (heads * L * L, L * L, L, 1), (*tok, E, 1), (L * E, hd, E))
        ctx.save_for_backward(q, k, v, P, o)
        ctx.heads = heads
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, P, o = ctx.saved_tensors
        heads = ctx.heads
        do = _chk(do)
        N, L, E = q.shape
        hd = E // heads
        tok = (L * E, hd)
        pb = (heads * L * L, L * L)
        scale = 1.0 / math.sqrt(hd)
        dv = torch.empty_like(v)
        gemm(P, do, dv, L, hd, L, N, heads, (*pb, 1, L), (*tok, E, 1), (L * E, hd, E))            # dV = P^T dO
        dP = torch.empty_like(P)
        vc = _center_tokens(v) if v.dtype == torch.float32 else v   # dS is invariant under a common shift of the values
        if _softmax_gemm(do, vc, P, dP, L, hd, N, heads, (*tok, E, 1), (*tok, 1, E), (*pb, L), scale):             # dS out of the dP product
            pass
        elif not _dsoftmax_gemm(do, vc, P, o, dP, L, hd, N, heads, (*tok, E, 1), (*tok, 1, E), (*pb, L), scale):   # opt-in variant with D = rowsum(dO * O)
            gemm(do, vc, dP, L, L, hd, N, heads, (*tok, E, 1), (*tok, 1, E), (*pb, L))              # dP = dO V^T
            lib.call("stc_softmax_rows_bwd", P, dP, dP, N * heads * L, L, scale, dtype_code(q.dtype), stream_ptr())
        dq = torch.empty_like(q)
        gemm(dP, k, dq, L, hd, L, N, heads, (*pb, L, 1), (*tok, E, 1), (L * E, hd, E))             # dQ = dS K
        dk = torch.empty_like(k)
        gemm(dP, q, dk, L, hd, L, N, heads, (*pb, 1, L), (*tok, E, 1), (L * E, hd, E))             # dK = dS^T Q
        return dq, dk, dv, None


def attention(q, k, v, heads: int):
    return _Attention.apply(q, k, v, heads)


class _InProj(Function):
    """nn.MultiheadAttention's packed in-projection: (q,k,v) -> (q W_q^T + b_q, k W_k^T + b_k, v W_v^T + b_v)
    with in_proj_weight (3E, E) / in_proj_bias (3E); the parameter gradients are produced whole."""

    @staticmethod
    def forward(ctx, q, k, v, W, b, pobjs):
        q, k, v = _chk(q), _chk(k), _chk(v)
        N, L, E = q.shape
        rows = N * L
        outs = []
        for i, t in enumerate((q, k, v)):
            wp = pack_weight(W[i * E:(i + 1) * E].view(E, E, 1, 1), t.dtype)
            outs.append(conv_fprop(t.view(1, 1, rows, E), wp, b[i * E:(i + 1) * E], None, E, 1, 1).view(N, L, E))
        ctx.save_for_backward(q, k, v, W)
        ctx.pobjs = pobjs
        return tuple(outs)

    @staticmethod
    def backward(ctx, dq, dk, dv):
        q, k, v, W = ctx.saved_tensors
        N, L, E = q.shape
        rows = N * L
        dev = q.device
        dW = _grad_buf(ctx.pobjs[0], W.shape, dev)
        db = _grad_buf(ctx.pobjs[1], (3 * E,), dev)
        dins = []
        for i, (t, g) in enumerate(((q, dq), (k, dk), (v, dv))):
            g = _chk(g)
            colsum(rows, E, g, db[i * E:(i + 1) * E])
            conv_wgrad(t.view(1, 1, rows, E), g.view(1, 1, rows, E), 1, 1, dW[i * E:(i + 1) * E].view(E, E, 1, 1))
            wpt = pack_weight(W[i * E:(i + 1) * E].view(E, E, 1, 1), g.dtype, transpose_flip=True)
            dins.append(conv_fprop(g.view(1, 1, rows, E), wpt, None, None, E, 1, 1).view(N, L, E))
        return dins[0], dins[1], dins[2], dW, db, None


def _gemm_f32out(A, B, C, M, N, K, batch, sA, sB, sC):
    """`batch` products of bf16 A(m,k), B(k,n) -> fp32 C on the tcgen05 engine (stc_gemm_f32out); element strides
    sA = (batch, m, k), sB = (batch, k, n), sC = (batch, m)."""
    d = GemmDesc(M, N, K, batch, 1, sA[0], 0, sA[1], sA[2], sB[0], 0, sB[1], sB[2], sC[0], 0, sC[1], 1.0, 0.0)
    _dense("gemm", 2.0 * M * N * K * batch, lambda: lib.call("stc_gemm_f32out", A, B, C, d, stream_ptr()))


def _uniform_stride(ts) -> Optional[int]:
    """Element stride between consecutive tensors when they are equally spaced in memory (16-byte granular), else None.
    Inside a Trainer step the bf16 weight packs (StepCache arena) and the gradient buffers (GradArena) of q/k/v are."""
    if len(ts) < 2:
        return 0
    d = ts[1].data_ptr() - ts[0].data_ptr()
    if d <= 0 or d % 16 or any(b.data_ptr() - a.data_ptr() != d for a, b in zip(ts[1:], ts[2:])):
        return None
    return d // ts[0].element_size()


class _FusedLinearPairs(Function):
    """J chained Linear pairs with nothing in between, folded into ONE token GEMM:

        y[:, j*Eo:(j+1)*Eo] = (x W1_j^T) W2_j^T + b2_j  =  x (W2_j W1_j)^T + b2_j        (+ residual when J == 1)

    TransformerLayer (unet_backbone.py:199-208) applies q/k/v (bias-free Linear) and then nn.MultiheadAttention's
    in-projection, and fc1 then fc2, without any nonlinearity: q/k/v + in_proj are J = 3 pairs sharing x (one GEMM with
    3E output columns, one dgrad GEMM that also sums the three input gradients, one wgrad GEMM), fc2(fc1(x)) + x is J = 1.
    The folded weights W2_j W1_j (E x E, bf16 from the bf16 operands the unfolded path would use) cost four E^3 products per
    pair and step; the parameter gradients follow from G_j = dy_j^T x:  dW2_j = G_j W1_j^T,  dW1_j = W2_j^T G_j.
    W2 is passed whole ((J*Eo, Em), e.g. in_proj_weight) so its gradient is produced whole."""

    @staticmethod
    def forward(ctx, x, residual, W2, b2, pobjs, *W1s):
        x = _chk(x)
        J = len(W1s)
        Em, Ei = W1s[0].shape
        Eo = W2.shape[0] // J
        rows = x.numel() // Ei
        dev, dt = x.device, x.dtype
        w2b = pack_weight(W2.view(J * Eo, Em, 1, 1), dt).view(J * Eo, Em)          # bf16 copies (batched by the StepCache)
        w1b = [pack_weight(w.view(Em, Ei, 1, 1), dt).view(Em, Ei) for w in W1s]
        weff = torch.empty((J * Eo, Ei), dtype=dt, device=dev)                       # fprop operand  [Cout][Cin]
        wefft = torch.empty((Ei, J * Eo), dtype=dt, device=dev)                      # dgrad operand  [Cin][Cout]
        s1 = _uniform_stride(w1b)
        # Weff_j[o][i] = sum_m W2_j[o][m] W1_j[m][i];  WeffT[i][j*Eo + o] = the same, written transposed (A = W1_j^T, B = W2_j^T)
        if J > 1 and s1 is not None:      # the J products as ONE batched launch each
            gemm(w2b, w1b[0], weff, Eo, Ei, Em, J, 1, (Eo * Em, 0, Em, 1), (s1, 0, Ei, 1), (Eo * Ei, 0, Ei))
            gemm(w1b[0], w2b, wefft, Ei, Eo, Em, J, 1, (s1, 0, 1, Ei), (Eo * Em, 0, 1, Em), (Eo, 0, J * Eo))
        else:
            for j in range(J):
                gemm(w2b[j * Eo:], w1b[j], weff[j * Eo:], Eo, Ei, Em, 1, 1, (0, 0, Em, 1), (0, 0, Ei, 1), (0, 0, Ei))
                gemm(w1b[j], w2b[j * Eo:], wefft[:, j * Eo:], Ei, Eo, Em, 1, 1, (0, 0, 1, Ei), (0, 0, 1, Em), (0, 0, J * Eo))
        if residual is not None:
            residual = _chk(residual).view(1, 1, rows, J * Eo)
        y = conv_fprop(x.view(1, 1, rows, Ei), weff.view(1, J * Eo, Ei), b2, residual, J * Eo, 1, 1)
        ctx.save_for_backward(x, W2, wefft, w2b, *w1b)
        ctx.meta = (J, Eo, Em, Ei, rows, pobjs, b2 is not None, residual is not None)
        return y.view(*x.shape[:-1], J * Eo)

    @staticmethod
    def backward(ctx, dy):
        x, W2, wefft, w2b, *w1b = ctx.saved_tensors
        J, Eo, Em, Ei, rows, pobjs, has_bias, has_res = ctx.meta
        pW2, pb2, pW1s = pobjs
        dy = _chk(dy)
        dev = dy.device
        db2 = colsum(rows, J * Eo, dy, _grad_buf(pb2, (J * Eo,), dev)) if has_bias else None
        # G^T (Ei x J*Eo, fp32) = x^T dy: the ordinary 1x1 wgrad of the folded layer
        gt = _wgrad_ws(Ei * J * Eo, dev)
        _dense("conv_wgrad", 2.0 * rows * Ei * J * Eo,
               lambda: lib.call("stc_conv_wgrad", x, dy, gt, 1, 1, rows, Ei, J * Eo, 1, 1, dtype_code(x.dtype), config.engine, stream_ptr()))
        dx = None
        if ctx.needs_input_grad[0]:
            dx = conv_fprop(dy.view(1, 1, rows, J * Eo), wefft.view(1, Ei, J * Eo), None, None, Ei, 1, 1).view(x.shape)
        gtb = pack_weight(gt.view(Ei, J * Eo, 1, 1), x.dtype, cache=False).view(Ei, J * Eo)     # bf16 copy of G^T
        dW2 = _grad_buf(pW2, (J * Eo, Em), dev)
        dW1s = [_grad_buf(pW1s[j], (Em, Ei), dev) for j in range(J)]
        s1, sg = _uniform_stride(w1b), _uniform_stride(dW1s)
        # dW2_j[o][m] = sum_i G_j[o][i] W1_j[m][i]     (A = G_j read transposed from G^T, B = W1_j^T)
        # dW1_j[m][i] = sum_o W2_j[o][m] G_j[o][i]     (A = W2_j^T, B = G_j)
        if J > 1 and s1 is not None and sg is not None:
            _gemm_f32out(gtb, w1b[0], dW2, Eo, Em, Ei, J, (Eo, 1, J * Eo), (s1, 1, Ei), (Eo * Em, Em))
            _gemm_f32out(w2b, gtb, dW1s[0], Em, Ei, Eo, J, (Eo * Em, 1, Em), (Eo, 1, J * Eo), (sg, Ei))
        else:
            for j in range(J):
                _gemm_f32out(gtb[:, j * Eo:], w1b[j], dW2[j * Eo:], Eo, Em, Ei, 1, (0, 1, J * Eo), (0, 1, Ei), (0, Em))
                _gemm_f32out(w2b[j * Eo:], gtb[:, j * Eo:], dW1s[j], Em, Ei, Eo, 1, (0, 1, Em), (0, 1, J * Eo), (0, Ei))
        dres = dy if (has_res and ctx.needs_input_grad[1]) else None
        return (dx, dres, dW2, db2, None, *dW1s)


def fused_linear_pairs(x, W1s, W2, b2=None, residual=None):
    """x (..., Ei) -> (..., J*Eo): see _FusedLinearPairs.  W1s: J Linear weights (Em, Ei); W2: (J*Eo, Em); b2: (J*Eo) or None."""
    return _FusedLinearPairs.apply(x, residual, W2, b2, (W2, b2, tuple(W1s)), *W1s)


class _AttentionPacked(Function):
    """_Attention on a packed (N, L, 3E) projection [q | k | v] (the output of the folded q/k/v + in_proj GEMM): the batched
    GEMMs read the three column blocks in place by stride and write dq/dk/dv into one packed gradient."""

    @staticmethod
    def forward(ctx, qkv, heads: int):
        qkv = _chk(qkv)
        N, L, E3 = qkv.shape
        E = E3 // 3
        hd = E // heads
        dev = qkv.device
        q, k, v = qkv[..., :E], qkv[..., E:2 * E], qkv[..., 2 * E:]
        P = torch.empty((N, heads, L, L), dtype=qkv.dtype, device=dev)
        pk = (L * E3, hd)   # batch strides of a packed (N, L, 3E) tensor split into heads
        scale = 1.0 / math.sqrt(hd)
        if not _softmax_gemm(q, k, P, None, L, hd, N, heads, (*pk, E3, 1), (*pk, 1, E3), (heads * L * L, L * L, L), scale):
            gemm(q, k, P, L, L, hd, N, heads, (*pk, E3, 1), (*pk, 1, E3), (heads * L * L, L * L, L))
            lib.call("stc_softmax_rows_fwd", P, P, N * heads * L, L, scale, dtype_code(qkv.dtype), stream_ptr())
        o = torch.empty((N, L, E), dtype=qkv.dtype, device=dev)
        gemm(P, v, o, L, hd, L, N, heads, (heads * L * L, L * L, L, 1), (*pk, E3, 1), (L * E, hd, E))
        ctx.save_for_backward(qkv, P, o)
        ctx.heads = heads
        return o

    @staticmethod
    def backward(ctx, do):
        qkv, P, o = ctx.saved_tensors
        heads = ctx.heads
        do = _chk(do)
        N, L, E3 = qkv.shape
        E = E3 // 3
        hd = E // heads
        q, k, v = qkv[..., :E], qkv[..., E:2 * E], qkv[..., 2 * E:]
        tok = (L * E, hd)
        pk = (L * E3, hd)
        pb = (heads * L * L, L * L)
        scale = 1.0 / math.sqrt(hd)
        dqkv = torch.empty_like(qkv)
        dq, dk, dv = dqkv[..., :E], dqkv[..., E:2 * E], dqkv[..., 2 * E:]
        gemm(P, do, dv, L, hd, L, N, heads, (*pb, 1, L), (*tok, E, 1), (*pk, E3))                   # dV = P^T dO
        dP = torch.empty_like(P)
        if _softmax_gemm(do, v, P, dP, L, hd, N, heads, (*tok, E, 1), (*pk, 1, E3), (*pb, L), scale):              # dS out of the dP product
            pass
        elif not _dsoftmax_gemm(do, v, P, o, dP, L, hd, N, heads, (*tok, E, 1), (*pk, 1, E3), (*pb, L), scale):   # opt-in variant with D = rowsum(dO * O)
            gemm(do, v, dP, L, L, hd, N, heads, (*tok, E, 1), (*pk, 1, E3), (*pb, L))               # dP = dO V^T
            lib.call("stc_softmax_rows_bwd", P, dP, dP, N * heads * L, L, scale, dtype_code(qkv.dtype), stream_ptr())
        gemm(dP, k, dq, L, hd, L, N, heads, (*pb, L, 1), (*pk, E3, 1), (*pk, E3))                   # dQ = dS K
        gemm(dP, q, dk, L, hd, L, N, heads, (*pb, 1, L), (*pk, E3, 1), (*pk, E3))                   # dK = dS^T Q
        return dqkv, None


def attention_packed(qkv, heads: int):
    return _AttentionPacked.apply(qkv, heads)


def in_proj(q, k, v, mha: torch.nn.MultiheadAttention):
    return _InProj.apply(q, k, v, mha.in_proj_weight, mha.in_proj_bias, (mha.in_proj_weight, mha.in_proj_bias))


# ---------------------------------------------------------------------------------------------
# classifier + loss
# ---------------------------------------------------------------------------------------------
class _ClsSeg(Function):
    """Dropout2d mask (optional, (N,Cin) fp32 already scaled) + Conv2d(Cin, Ccls, 1): NHWC -> NCHW fp32 logits."""

    @staticmethod
    def forward(ctx, x, weight, bias, mask, pobjs):
        x = _chk(x)
        N, H, W, Cin = x.shape
        Ccls = weight.shape[0]
        code = dtype_code(x.dtype)
        if mask is not None:
            mask = _chk(mask.float())
        logits = torch.empty((N, Ccls, H, W), dtype=torch.float32, device=x.device)
        lib.call("stc_cls_fwd", x, weight.view(Ccls, Cin), bias, mask, logits, N, H * W, Cin, Ccls, code, stream_ptr())
        ctx.save_for_backward(x, weight, mask)
        ctx.pobjs = pobjs
        return logits

    @staticmethod
    def backward(ctx, dl):
        xm, weight, mask = ctx.saved_tensors
        dl = _chk(dl)
        N, H, W, Cin = xm.shape
        Ccls = weight.shape[0]
        code = dtype_code(xm.dtype)
        dev = xm.device
        dx = torch.empty_like(xm) if ctx.needs_input_grad[0] else None
        dW = _grad_buf(ctx.pobjs[0], weight.shape, dev, zero=True)
        db = _grad_buf(ctx.pobjs[1], (Ccls,), dev, zero=True)
        lib.call("stc_cls_bwd", dl, xm, weight.view(Ccls, Cin), mask, dx, dW, db, N, H * W, Cin, Ccls, None, 0, code, stream_ptr())
        return dx, dW, db, None, None


def cls_seg(x, conv_seg: torch.nn.Conv2d, mask=None):
    return _ClsSeg.apply(x, conv_seg.weight, conv_seg.bias, mask, (conv_seg.weight, conv_seg.bias))


class _SegLoss(Function):
    """(loss_ce_mean_over_all_pixels, loss_dice, acc_seg) from NCHW fp32 logits and (N,H,W) int64 labels."""

    @staticmethod
    def forward(ctx, logits, label, ignore_index: int, smooth: float):
        logits = _chk(logits)
        label = _chk(label)
        N, C, H, W = logits.shape
        n = lib.raw("stc_seg_loss_stats_len")(N, C)
        stats = torch.empty(n, dtype=torch.float64, device=logits.device)
        out3 = torch.empty(3, dtype=torch.float32, device=logits.device)
        lib.call("stc_seg_loss_fwd", logits, label, stats, out3, N, H * W, C, ignore_index, float(smooth), stream_ptr())
        ctx.save_for_backward(logits, label, stats)
        ctx.meta = (ignore_index, float(smooth))
        return out3[0], out3[1], out3[2]

    @staticmethod
    def backward(ctx, g_ce, g_dice, _g_acc):
        logits, label, stats = ctx.saved_tensors
        ignore_index, smooth = ctx.meta
        N, C, H, W = logits.shape
        dl = torch.empty_like(logits)
        g_ce = None if g_ce is None else _chk(g_ce.float())
        g_dice = None if g_dice is None else _chk(g_dice.float())
        lib.call("stc_seg_loss_bwd", logits, label, stats, g_ce, g_dice, dl, N, H * W, C, ignore_index, smooth, stream_ptr())
        return dl, None, None, None


def seg_loss(logits, label, ignore_index=255, smooth=1.0):
    return _SegLoss.apply(logits, labels_to_int64(label), ignore_index, smooth)


# ---------------------------------------------------------------------------------------------
# layout conversion and inference / metric helpers (no autograd)
# ---------------------------------------------------------------------------------------------
_NORM_CACHE = {}


def _norm_vectors(device, C, mean, std):
    key = (device.type, device.index, C, tuple(mean), tuple(std))
    hit = _NORM_CACHE.get(key)
    if hit is None:
        m = torch.tensor([float(mean[i % len(mean)]) for i in range(C)], dtype=torch.float32, device=device)
        r = torch.tensor([1.0 / float(std[i % len(std)]) for i in range(C)], dtype=torch.float32, device=device)
        hit = _NORM_CACHE[key] = (m, r)
    return hit


class NHWCImage(torch.Tensor):
    """Already-normalised (N, H, W, C) activations (the output of augment_batch_u8).  The layout travels with the TYPE: a tensor
    subclass survives clone() / to() / copy_() - which Trainer.capture / step_graph apply to their inputs - where a Python attribute
    on the tensor object would be dropped silently and the NHWC data re-read as NCHW."""

    @staticmethod
    def wrap(t: torch.Tensor) -> "NHWCImage":
        return t.as_subclass(NHWCImage)


def image_to_nhwc(img: torch.Tensor, dtype: torch.dtype, norm_cfg: Optional[dict] = None) -> torch.Tensor:
    """Module input -> NHWC activations.  float (N,C,H,W): the reference's interface (already normalised by its CPU pipeline).
    uint8 (N,H,W,C): decoded pixels straight from the loader; `norm_cfg = dict(mean, std, to_rgb)` (the config's img_norm_cfg,
    my_config/STC-UNet.py:35) is applied on the device."""
    if isinstance(img, NHWCImage):           # already normalised NHWC activations (augment_batch_u8)
        img = img.as_subclass(torch.Tensor)
        if img.dim() != 4:
            raise ValueError(f"NHWCImage must be (N, H, W, C), got {tuple(img.shape)}")
        return _chk(img) if img.dtype == dtype else _chk(img.to(dtype))
    if img.dtype == torch.uint8:
        img = _chk(img)
        if img.dim() != 4 or img.shape[-1] > 4:
            raise ValueError(f"uint8 images must be (N, H, W, C<=4) as decoded, got {tuple(img.shape)}")
        N, H, W, C = img.shape
        cfg = norm_cfg or {}
        mean, inv_std = _norm_vectors(img.device, C, cfg.get("mean", [0.0]), cfg.get("std", [1.0]))
        out = torch.empty((N, H, W, C), dtype=dtype, device=img.device)
        lib.call("stc_image_u8_to_nhwc", img, out, mean, inv_std, N * H * W, C, C, int(bool(cfg.get("to_rgb", False))), dtype_code(dtype),
                 stream_ptr())
        return out
    img = _chk(img.float())
    N, C, H, W = img.shape
    out = torch.empty((N, H, W, C), dtype=dtype, device=img.device)
    lib.call("stc_nchw_to_nhwc", img, out, N, C, H, W, C, dtype_code(dtype), stream_ptr())
    return out


def draw_crop_flip(n: int, src_hw, crop_hw, flip_prob: float = 0.5, rng=None) -> torch.Tensor:
    """(n, 3) int32 {y0, x0, flip} drawn on the host the way the reference pipeline does it per image: RandomCrop.get_crop_bbox
    (transforms.py:599-608: offsets uniform in [0, margin]) and RandomFlip (`np.random.rand() < prob`, :358-360)."""
    import numpy as np
    rng = rng or np.random
    mh, mw = max(src_hw[0] - crop_hw[0], 0), max(src_hw[1] - crop_hw[1], 0)
    g = np.empty((n, 3), dtype=np.int32)
    for i in range(n):
        g[i, 0] = rng.randint(0, mh + 1)
        g[i, 1] = rng.randint(0, mw + 1)
        g[i, 2] = 1 if rng.rand() < flip_prob else 0
    return torch.from_numpy(g)


def augment_batch_u8(img_u8: torch.Tensor, label_u8: Optional[torch.Tensor], geom: torch.Tensor, crop_hw, dtype: torch.dtype,
                     norm_cfg: Optional[dict] = None, pad_val: float = 0.0, seg_pad_val: int = 255):
    """RandomCrop -> RandomFlip -> Normalize -> Pad of the training pipeline on the device.  img_u8 (N, Hs, Ws, C) uint8,
    label_u8 (N, Hs, Ws) uint8 or None, geom (N, 3) int32 from draw_crop_flip.  Returns NHWC activations (N, H, W, C) in `dtype` - which the
    backbone takes as they are - and int64 labels (N, 1, H, W)."""
    img_u8 = _chk(img_u8)
    N, Hs, Ws, C = img_u8.shape
    H, W = crop_hw
    geom = geom.to(device=img_u8.device, dtype=torch.int32).contiguous()
    cfg = norm_cfg or {}
    mean, inv_std = _norm_vectors(img_u8.device, C, cfg.get("mean", [0.0]), cfg.get("std", [1.0]))
    out = torch.empty((N, H, W, C), dtype=dtype, device=img_u8.device)
    lib.call("stc_image_u8_crop_flip_to_nhwc", img_u8, out, geom, mean, inv_std, N, Hs, Ws, H, W, C, C, int(bool(cfg.get("to_rgb", False))),
             float(pad_val), dtype_code(dtype), stream_ptr())
    out = NHWCImage.wrap(out)   # image_to_nhwc (the backbones' first step) passes it through
    lab = None
    if label_u8 is not None:
        label_u8 = _chk(label_u8)
        lab = torch.empty((N, 1, H, W), dtype=torch.int64, device=img_u8.device)
        lib.call("stc_label_u8_crop_flip_i64", label_u8, lab, geom, N, Hs, Ws, H, W, int(seg_pad_val), stream_ptr())
    return out, lab


def labels_to_int64(label: torch.Tensor) -> torch.Tensor:
    """8-bit label maps (annotation PNGs) are widened on the device; int64 passes through."""
    if label.dtype == torch.int64:
        return label
    if label.dtype != torch.uint8:
        raise TypeError(f"labels must be int64 or uint8, got {label.dtype}")
    label = _chk(label)
    out = torch.empty(label.shape, dtype=torch.int64, device=label.device)
    lib.call("stc_widen_u8_i64", label, out, label.numel(), stream_ptr())
    return out


def argmax_nchw(preds: torch.Tensor, count: Optional[torch.Tensor] = None) -> torch.Tensor:
    preds = _chk(preds)
    N, C, H, W = preds.shape
    out = torch.empty((N, H, W), dtype=torch.int64, device=preds.device)
    lib.call("stc_argmax", preds, count, out, N, C, H * W, stream_ptr())
    return out


def slide_accum(crop, preds, count, y1, x1):
    crop = _chk(crop)
    N, C, H, W = preds.shape
    lib.call("stc_slide_accum", crop, preds, count, N, C, H, W, crop.shape[2], crop.shape[3], y1, x1, stream_ptr())


def confusion_hist(pred: torch.Tensor, label: torch.Tensor, num_classes: int, ignore_index: int = 255,
                   cm: Optional[torch.Tensor] = None, areas: Optional[torch.Tensor] = None):
    """Accumulates the int64 confusion matrix CM[label, pred] and the four area vectors of
    intersect_and_union (metrics.py:75-87) on the device."""
    pred = _chk(pred.to(torch.int64))
    if label.dtype not in (torch.uint8, torch.int64):
        label = label.to(torch.int64)
    label = _chk(label)
    dev = pred.device
    if cm is None:
        cm = torch.zeros((num_classes, num_classes), dtype=torch.int64, device=dev)
    if areas is None:
        areas = torch.zeros((4, num_classes), dtype=torch.int64, device=dev)
    lib.call("stc_confusion_hist", pred, label, int(label.dtype == torch.uint8), pred.numel(), num_classes, ignore_index,
             cm, areas, stream_ptr())
    return cm, areas


# ---------------------------------------------------------------------------------------------
# small differentiable helpers
# ---------------------------------------------------------------------------------------------
class _Add(Function):
    @staticmethod
    def forward(ctx, a, b):
        return add(_chk(a), _chk(b))

    @staticmethod
    def backward(ctx, g):
        return g, g


def add_autograd(a, b):
    return _Add.apply(a, b)


def _copy_rows(src, dst, src_off, dst_off, count):
    N, rs, C = src.shape[0], src.shape[1], src.shape[-1]
    lib.call("stc_copy_rows", src, dst, N, rs, dst.shape[1], C, src_off, dst_off, count, dtype_code(src.dtype), stream_ptr())


class _SliceRows(Function):
    """(N, R, 1, C) -> (N, count, 1, C) starting at row `off`."""

    @staticmethod
    def forward(ctx, x, off, count):
        x = _chk(x)
        out = torch.empty((x.shape[0], count, 1, x.shape[-1]), dtype=x.dtype, device=x.device)
        _copy_rows(x, out, off, 0, count)
        ctx.meta = (x.shape, off, count)
        return out

    @staticmethod
    def backward(ctx, g):
        shape, off, count = ctx.meta
        g = _chk(g)
        dx = torch.zeros(shape, dtype=g.dtype, device=g.device)
        _copy_rows(g, dx, 0, off, count)
        return dx, None, None


def slice_rows(x, off, count):
    return _SliceRows.apply(x, off, count)


class _CatRows(Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = _chk(a), _chk(b)
        ra, rb = a.shape[1], b.shape[1]
        out = torch.empty((a.shape[0], ra + rb, 1, a.shape[-1]), dtype=a.dtype, device=a.device)
        _copy_rows(a, out, 0, 0, ra)
        _copy_rows(b, out, 0, ra, rb)
        ctx.meta = (ra, rb)
        return out

    @staticmethod
    def backward(ctx, g):
        ra, rb = ctx.meta
        g = _chk(g)
        N, C = g.shape[0], g.shape[-1]
        ga = torch.empty((N, ra, 1, C), dtype=g.dtype, device=g.device)
        gb = torch.empty((N, rb, 1, C), dtype=g.dtype, device=g.device)
        _copy_rows(g, ga, 0, 0, ra)
        _copy_rows(g, gb, ra, 0, rb)
        return ga, gb


def cat_rows(a, b):
    return _CatRows.apply(a, b)


class _Scale(Function):
    @staticmethod
    def forward(ctx, x, alpha):
        ctx.alpha = alpha
        x = _chk(x.float())
        out = torch.zeros_like(x)
        lib.call("stc_axpy_f32", x, out, float(alpha), x.numel(), stream_ptr())
        return out

    @staticmethod
    def backward(ctx, g):
        g = _chk(g.float())
        out = torch.zeros_like(g)
        lib.call("stc_axpy_f32", g, out, float(ctx.alpha), g.numel(), stream_ptr())
        return out, None


def scale(x, alpha):
    return x if float(alpha) == 1.0 else _Scale.apply(x, alpha)


# ---------------------------------------------------------------------------------------------
# family B helpers: stand-alone bilinear x2 (InterpConv) and channel concat (UpConvBlock)
# ---------------------------------------------------------------------------------------------
class _Upsample2x(Function):
    @staticmethod
    def forward(ctx, low, align_corners: bool):
        low = _chk(low)
        N, h, w, C = low.shape
        out = torch.empty((N, 2 * h, 2 * w, C), dtype=low.dtype, device=low.device)
        lib.call("stc_upcat_fwd", None, low, out, N, 2 * h, 2 * w, 0, h, w, C, int(align_corners), dtype_code(low.dtype), stream_ptr())
        ctx.meta = (N, h, w, C, int(align_corners))
        return out

    @staticmethod
    def backward(ctx, dout):
        N, h, w, C, ac = ctx.meta
        dout = _chk(dout)
        dlow = torch.empty((N, h, w, C), dtype=dout.dtype, device=dout.device)
        lib.call("stc_upcat_bwd", dout, None, dlow, N, 2 * h, 2 * w, 0, h, w, C, ac, dtype_code(dout.dtype), stream_ptr())
        return dlow, None


def upsample2x(x, align_corners=False):
    return _Upsample2x.apply(x, align_corners)


class _ConcatChannels(Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = _chk(a), _chk(b)
        Ca, Cb = a.shape[-1], b.shape[-1]
        out = torch.empty((*a.shape[:-1], Ca + Cb), dtype=a.dtype, device=a.device)
        lib.call("stc_concat_channels", a, b, out, a.numel() // Ca, Ca, Cb, dtype_code(a.dtype), stream_ptr())
        ctx.meta = (Ca, Cb)
        return out

    @staticmethod
    def backward(ctx, g):
        Ca, Cb = ctx.meta
        g = _chk(g)
        P = g.numel() // (Ca + Cb)
        ga = torch.empty((*g.shape[:-1], Ca), dtype=g.dtype, device=g.device) if ctx.needs_input_grad[0] else None
        gb = torch.empty((*g.shape[:-1], Cb), dtype=g.dtype, device=g.device) if ctx.needs_input_grad[1] else None
        lib.call("stc_split_channels", g, ga, gb, P, Ca, Cb, dtype_code(g.dtype), stream_ptr())
        return ga, gb


def concat_channels(a, b):
    return _ConcatChannels.apply(a, b)


class _CatN(Function):
    """cat([x0', x1, ..]) along channels; x0' = nearest x2 upsample of x0 when up0 (UNet++ DecoderBlock)."""

    @staticmethod
    def forward(ctx, up0: bool, *xs):
        xs = [_chk(x) for x in xs]
        assert 1 <= len(xs) <= 5
        N, h, w, _ = xs[0].shape
        H, W = (2 * h, 2 * w) if up0 else (h, w)
        chs = [x.shape[-1] for x in xs]
        out = torch.empty((N, H, W, sum(chs)), dtype=xs[0].dtype, device=xs[0].device)
        pad = [None] * (5 - len(xs))
        lib.call("stc_catn_fwd", *xs, *pad, *chs, *([0] * (5 - len(xs))), out, N, H, W, int(up0), dtype_code(out.dtype), stream_ptr())
        ctx.meta = (N, H, W, chs, int(up0), [tuple(x.shape) for x in xs])
        return out

    @staticmethod
    def backward(ctx, g):
        N, H, W, chs, up0, shapes = ctx.meta
        g = _chk(g)
        ds = [torch.empty(shp, dtype=g.dtype, device=g.device) if ctx.needs_input_grad[i + 1] else None for i, shp in enumerate(shapes)]
        pad = [None] * (5 - len(ds))
        lib.call("stc_catn_bwd", g, *ds, *pad, *chs, *([0] * (5 - len(chs))), N, H, W, up0, dtype_code(g.dtype), stream_ptr())
        return (None, *ds)


def cat_channels_n(xs, up0=False):
    return _CatN.apply(up0, *xs)


# ---------------------------------------------------------------------------------------------
# DeconvModule (unet.py:89-147): ConvTranspose2d(4, 2, 1) = 3x3 conv to 4*Cout sub-pixel channels + pixel shuffle, then BN + act
# ---------------------------------------------------------------------------------------------
class _DepthToSpace2(Function):
    @staticmethod
    def forward(ctx, z):
        z = _chk(z)
        N, H, W, C4 = z.shape
        out = torch.empty((N, 2 * H, 2 * W, C4 // 4), dtype=z.dtype, device=z.device)
        lib.call("stc_depth_to_space2", z, out, N, H, W, C4 // 4, 0, dtype_code(z.dtype), stream_ptr())
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = _chk(dout)
        N, H2, W2, C = dout.shape
        dz = torch.empty((N, H2 // 2, W2 // 2, 4 * C), dtype=dout.dtype, device=dout.device)
        lib.call("stc_depth_to_space2", dout, dz, N, H2 // 2, W2 // 2, C, 1, dtype_code(dout.dtype), stream_ptr())
        return dz


def depth_to_space2(z):
    return _DepthToSpace2.apply(z)


class _BnAct(Function):
    """Stand-alone BatchNorm + activation on an NHWC tensor (the norm after DeconvModule's pixel shuffle)."""

    @staticmethod
    def forward(ctx, y, gamma, beta, bn: BNState, act: int, pobjs):
        y = _chk(y)
        C = y.shape[-1]
        P = y.numel() // C
        mean, invstd, count = _bn_forward_stats(y, P, C, bn)
        a = torch.empty_like(y)
        lib.call("stc_bn_apply", y, mean, invstd, gamma, beta, a, P, C, act, dtype_code(y.dtype), stream_ptr())
        ctx.save_for_backward(y, gamma, beta, mean, invstd)
        ctx.meta = (act, bn, count, pobjs)
        return a

    @staticmethod
    def backward(ctx, da):
        y, gamma, beta, mean, invstd = ctx.saved_tensors
        act, bn, count, (pg, pb) = ctx.meta
        C = y.shape[-1]
        dy, dgamma, dbeta = _bn_backward(y, _chk(da), mean, invstd, gamma, beta, y.numel() // C, C, act, bn, count, pg, pb)
        return dy, dgamma, dbeta, None, None, None


def bn_act(y, bn: torch.nn.modules.batchnorm._BatchNorm, act: int, training: bool):
    sync = isinstance(bn, torch.nn.SyncBatchNorm)
    state = BNState(bn.running_mean, bn.running_var, bn.num_batches_tracked, 0.1 if bn.momentum is None else bn.momentum, bn.eps,
                    training or not bn.track_running_stats, sync=sync, group=getattr(bn, "process_group", None) if sync else None)
    return _BnAct.apply(y, bn.weight, bn.bias, state, act, (bn.weight, bn.bias))


_DECONV_TAPS = ((3, 1, 4), (4, 2, 0))   # [output parity][3x3 tap] -> 4x4 kernel index (4 = no contribution)


def deconv4x2_weight_as_conv3(weight: torch.Tensor) -> torch.Tensor:
    """ConvTranspose2d(k=4,s=2,p=1).weight (Cin,Cout,4,4) -> the equivalent 3x3 correlation weight (4*Cout, Cin, 3, 3) whose output
    channel (py*2+px)*Cout + co is output pixel (2m+py, 2l+px): oy = 2*iy - 1 + ky, so parity 0 reads rows m-1 (ky=3), m (ky=1) and
    parity 1 reads rows m (ky=2), m+1 (ky=0).  Plain differentiable index ops on the (small) weight: autograd maps the gradient back."""
    Cin, Cout = weight.shape[:2]
    idx = torch.tensor(_DECONV_TAPS, device=weight.device)                   # (2, 3)
    wp = torch.cat([weight, weight.new_zeros(Cin, Cout, 1, 4)], dim=2)       # ky = 4 -> zero row
    wp = torch.cat([wp, wp.new_zeros(Cin, Cout, 5, 1)], dim=3)               # kx = 4 -> zero column
    w = wp[:, :, idx][:, :, :, :, idx]                                       # (Cin, Cout, py, r, px, s)
    return w.permute(2, 4, 1, 0, 3, 5).reshape(4 * Cout, Cin, 3, 3).contiguous()


class _BiasIntoTrainBN(Function):
    """Identity on a conv bias that feeds a train-mode BN: its gradient is exactly zero (the batch mean removes it), so return
    zeros instead of the rounding noise a column sum of dy would give (same rule as _ConvBnAct)."""

    @staticmethod
    def forward(ctx, bias):
        ctx.pobj = bias
        return bias.view_as(bias)

    @staticmethod
    def backward(ctx, g):
        return _grad_buf(ctx.pobj, tuple(g.shape), g.device, zero=True)


def deconv4x2(x, weight, bias=None, bias_feeds_train_bn: bool = False):
    """ConvTranspose2d(Cin, Cout, kernel_size=4, stride=2, padding=1) on NHWC x: (N,H,W,Cin) -> (N,2H,2W,Cout)."""
    w3 = deconv4x2_weight_as_conv3(weight)
    if bias is not None and bias_feeds_train_bn:
        bias = _BiasIntoTrainBN.apply(bias)
    b4 = bias.repeat(4) if bias is not None else None
    return depth_to_space2(_Conv.apply(x, w3, b4, None, 0, (None, None, True)))
