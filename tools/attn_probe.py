"""Attention score products with the row softmax inside (stc_gemm_softmax / _bwd, two sweeps) against product + separate softmax pass; the
STC-UNet level-4 shape (32 batch-heads, 4096 tokens, head dim 256) and the level-5 one (1024 tokens)."""
import os, sys, math, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stc_unet_b200 as S
from stc_unet_b200 import ops
from stc_unet_b200._lib import GemmDesc, dtype_code, lib, stream_ptr
BF = torch.bfloat16; dev = torch.device("cuda:0")
def timeit(fn, name, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    print(f"{name:60s} {ms:8.3f} ms", flush=True)
    return ms
for L in (4096, 1024):
    hd, B = 256, 32
    scale = 1.0 / math.sqrt(hd)
    q = (torch.randn(B, L, hd, device=dev) * 0.5).to(BF); kk = (torch.randn(B, L, hd, device=dev) * 0.5).to(BF); v = torch.randn(B, L, hd, device=dev).to(BF)
    do = torch.randn(B, L, hd, device=dev).to(BF)
    P = torch.empty(B, L, L, device=dev, dtype=BF); P2 = torch.empty_like(P); dS = torch.empty_like(P); dS2 = torch.empty_like(P)
    sA, sB, sC = (L * hd, 0, hd, 1), (L * hd, 0, 1, hd), (L * L, 0, L)
    d = GemmDesc(L, L, hd, B, 1, *sA, *sB, *sC, 1.0, 0.0)
    def unfused_fwd():
        ops.gemm(q, kk, P, L, L, hd, B, 1, sA, sB, sC)
        lib.call("stc_softmax_rows_fwd", P, P, B * L, L, scale, dtype_code(BF), stream_ptr())
    def fused_fwd():
        lib.call("stc_gemm_softmax", q, kk, P2, d, scale, dtype_code(BF), 0, stream_ptr())
    def unfused_bwd():
        ops.gemm(do, v, dS, L, L, hd, B, 1, sA, sB, sC)
        lib.call("stc_softmax_rows_bwd", P, dS, dS, B * L, L, scale, dtype_code(BF), stream_ptr())
    def fused_bwd():
        lib.call("stc_gemm_softmax_bwd", do, v, P, dS2, d, scale, dtype_code(BF), 0, stream_ptr())
    a = timeit(unfused_fwd, f"L={L}: QK^T + softmax pass")
    b = timeit(fused_fwd, f"L={L}: stc_gemm_softmax (two sweeps)")
    unfused_fwd(); fused_fwd(); torch.cuda.synchronize()
    print(f"    P max abs diff {float((P.float() - P2.float()).abs().max()):.3e}  rel l2 {float((P.float() - P2.float()).norm() / P.float().norm()):.3e}  row sums {float(P2.float().sum(-1).mean()):.5f}")
    c = timeit(unfused_bwd, f"L={L}: dO V^T + softmax-backward pass")
    e = timeit(fused_bwd, f"L={L}: stc_gemm_softmax_bwd (two sweeps)")
    unfused_bwd(); fused_bwd(); torch.cuda.synchronize()
    print(f"    dS rel l2 {float((dS.float() - dS2.float()).norm() / dS.float().norm()):.3e}")
    print(f"    forward {a:.3f} -> {b:.3f} ms, backward {c:.3f} -> {e:.3f} ms")
    del P, P2, dS, dS2
