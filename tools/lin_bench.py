"""Times the GEMM-shaped launches of the transformer blocks (1x1 conv = Linear, attention GEMMs) with CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stc_unet_b200 as S
from stc_unet_b200 import ops
BF = torch.bfloat16; dev = torch.device("cuda:0")

def timeit(fn, flops, name, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    print(f"{name:44s} {ms:8.3f} ms {flops / ms / 1e9:8.0f} TF/s")

for (N, hw, ci, co, k) in [(16, 64, 512, 512, 1), (16, 32, 512, 512, 1), (16, 64, 512, 512, 3), (16, 64, 256, 512, 3), (16, 32, 512, 512, 3), (16, 64, 1024, 256, 3)]:
    x = torch.randn(N, hw, hw, ci, device=dev).to(BF)
    w = torch.randn(co, ci, k, k, device=dev) / (ci * k * k) ** 0.5
    wp = ops.pack_weight(w, BF)
    res = torch.randn(N, hw, hw, co, device=dev).to(BF)
    fl = 2.0 * N * hw * hw * ci * co * k * k
    timeit(lambda: ops.conv_fprop(x, wp, None, None, co, k, k), fl, f"conv {ci}->{co} k{k} @{hw} N{N}")
    if k == 1:
        timeit(lambda: ops.conv_fprop(x, wp, None, res, co, k, k), fl, f"conv {ci}->{co} k{k} @{hw} N{N} +residual")
for (N, hw, ci, co, k) in [(16, 64, 512, 512, 1), (16, 32, 512, 512, 1), (16, 64, 512, 512, 3), (16, 64, 256, 512, 3), (16, 32, 512, 512, 3), (16, 64, 1024, 256, 3)]:
    x = torch.randn(N, hw, hw, ci, device=dev).to(BF)
    dy = torch.randn(N, hw, hw, co, device=dev).to(BF)
    ws = torch.zeros(k * k * ci * co, device=dev)
    timeit(lambda: S._lib.lib.call("stc_conv_wgrad", x, dy, ws, N, hw, hw, ci, co, k, k, 1, 0, S._lib.stream_ptr()), 2.0 * N * hw * hw * ci * co * k * k,
           f"wgrad {ci}->{co} k{k} @{hw} N{N}")
# attention GEMMs: heads*N = 32 batches, L = 4096, hd = 256
L, hd, B = 4096, 256, 32
q = torch.randn(B, L, hd, device=dev).to(BF); kk = torch.randn(B, L, hd, device=dev).to(BF); v = torch.randn(B, L, hd, device=dev).to(BF)
sc = torch.empty(B, L, L, device=dev, dtype=BF); o = torch.empty(B, L, hd, device=dev, dtype=BF)
timeit(lambda: ops.gemm(q, kk, sc, L, L, hd, B, 1, (L * hd, 0, hd, 1), (L * hd, 0, 1, hd), (L * L, 0, L)), 2.0 * B * L * L * hd, "QK^T  (K-major x K-major) 32x4096x4096x256")
timeit(lambda: ops.gemm(sc, v, o, L, hd, L, B, 1, (L * L, 0, L, 1), (L * hd, 0, hd, 1), (L * hd, 0, hd)), 2.0 * B * L * L * hd, "PV    (K-major x N-major) 32x4096x256x4096")
