"""ncu: BW_MICRO_REPS=1 ncu --set full --clock-control none -k regex:"bn_|softmax|add_n|ksa_|cls_|colsum" -c 40 ... (bound the launch count: a
full replay of every repetition does not fit a few GPU-minutes).
Times individual bandwidth-bound C-ABI calls at their heaviest in-step shapes (STC-UNet, N=16, 512x512, bf16): CUDA events, 5 reps
after warm-up, a 512 MB write between reps so nothing is served from L2.  Prints ms and algorithmic GB/s per call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stc_unet_b200 as S
from stc_unet_b200._lib import lib, stream_ptr
BF = torch.bfloat16
dev = torch.device("cuda:0")
sp = stream_ptr
flush = torch.empty(256 << 20, dtype=torch.float32, device=dev)
only = os.environ.get("ONLY")


REPS = int(os.environ.get("BW_MICRO_REPS", "5"))     # under `ncu --set full` use BW_MICRO_REPS=1 (every launch is replayed ~40 times)


def timeit(name, gbytes, fn, reps=None):
    reps = reps or REPS
    if only and only not in name:
        return
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print("%-44s %8.3f ms %8.0f GB/s" % (name, t, gbytes / t * 1e3), flush=True)


N, HW, C = 16, 512 * 512, 64
P = N * HW
x = torch.randn(P, C, device=dev).to(BF)
d = torch.randn(P, C, device=dev).to(BF)
o = torch.empty_like(x)
GB = lambda *ts: sum(t.numel() * t.element_size() for t in ts) / 1e9

# classifier backward (64 -> 3 classes)
dl = torch.randn(N, 3, HW, device=dev)
W = torch.randn(3, C, device=dev); mask = torch.rand(N, C, device=dev)
dW = torch.zeros(3, C, device=dev); db = torch.zeros(3, device=dev)
timeit("cls_bwd dx only", GB(dl, o), lambda: lib.call("stc_cls_bwd", dl, x, W, mask, o, None, None, N, HW, C, 3, None, 0, 1, sp()))
timeit("cls_bwd dW only", GB(dl, x), lambda: lib.call("stc_cls_bwd", dl, x, W, mask, None, dW, db, N, HW, C, 3, None, 0, 1, sp()))
logits = torch.empty(N, 3, HW, device=dev)
timeit("cls_fwd", GB(x, logits), lambda: lib.call("stc_cls_fwd", x, W, db, mask, logits, N, HW, C, 3, 1, sp()))

# column sums (bias gradients): out_proj (65536 x 512) and folded qkv (65536 x 1536)
for rows, cols in ((65536, 512), (65536, 1536), (16384, 1536)):
    t = torch.randn(rows, cols, device=dev).to(BF); out = torch.zeros(cols, device=dev)
    timeit(f"colsum {rows}x{cols}", GB(t), lambda: lib.call("stc_colsum", t, out, rows, cols, 0, 1, sp()))

# BN pieces at level 1
mean = torch.zeros(C, device=dev); invstd = torch.ones(C, device=dev); gamma = torch.ones(C, device=dev); beta = torch.zeros(C, device=dev)
sums = torch.zeros(2 * C, dtype=torch.float64, device=dev)
ws = torch.empty(lib.raw("stc_bn_ws_bytes")(P, C), dtype=torch.uint8, device=dev)
ua = torch.rand(N, C, device=dev); ub = torch.rand(N, C, device=dev)
timeit("bn_reduce L1", GB(x), lambda: lib.call("stc_bn_reduce", x, sums, P, C, ws, ws.numel(), 1, sp()))
timeit("bn_apply L1", GB(x, o), lambda: lib.call("stc_bn_apply", x, mean, invstd, gamma, beta, o, P, C, 1, 1, sp()))
timeit("bn_bwd_reduce L1", GB(x, d), lambda: lib.call("stc_bn_bwd_reduce", x, d, mean, invstd, gamma, beta, sums, P, C, 1, ws, ws.numel(), 1, sp()))
timeit("bn_bwd_reduce_aff L1", GB(x, d), lambda: lib.call("stc_bn_bwd_reduce_aff", x, d, ua, ub, 1.0, HW, N, mean, invstd, gamma, beta, sums, C, ws, ws.numel(), 1, sp()))
timeit("bn_bwd_apply L1", GB(x, d, o), lambda: lib.call("stc_bn_bwd_apply", x, d, mean, invstd, gamma, beta, sums, float(P), o, P, C, 1, 0, 1, sp()))
timeit("bn_bwd_apply_aff L1", GB(x, d, o), lambda: lib.call("stc_bn_bwd_apply_aff", x, d, ua, ub, 1.0, HW, N, mean, invstd, gamma, beta, sums, float(P), o, C, 1, sp()))

# fused Up.forward pieces at up4 (64 + 64 channels at 512x512)
from stc_unet_b200 import ops
skip = torch.randn(N, 512, 512, 64, device=dev, dtype=BF, requires_grad=True)
low = torch.randn(N, 256, 256, 64, device=dev, dtype=BF, requires_grad=True)
if not only or "upcat" in only:
    prof = ops.LaunchProfiler(time_all=True) if hasattr(ops.LaunchProfiler, "__init__") else None
    names = {}
    orig = lib.call
    import types
    def timed_call(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = orig(name, *a); e1.record()
        names.setdefault(name, []).append((e0, e1))
        return r
    for it in range(3):
        if it == 2:
            lib.call = timed_call
        out = ops.upcat_coordatt(skip, low, True, lambda y, n, h, w: torch.sigmoid(y))
        out.backward(torch.ones_like(out))
        skip.grad = low.grad = None
    torch.cuda.synchronize()
    if "call" in lib.__dict__:
        del lib.__dict__["call"]
    for k, evs in names.items():
        print("%-44s %8.3f ms  (%d calls, up4 shape, warm L2)" % ("upcat:" + k, sum(a.elapsed_time(b) for a, b in evs), len(evs)))
