"""Timing probe of the short-K GEMMs (QK^T, token Linear) - used with STC_CTA2 (and, while it existed, a debug knob that skipped parts of the epilogue: profiles/r2_gemm_probe.txt) to see what bounds them."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
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
    print(f"{name:44s} {ms:8.3f} ms {flops / ms / 1e9:8.0f} TF/s", flush=True)
L, hd, B = 4096, 256, 32
q = torch.randn(B, L, hd, device=dev).to(BF); kk = torch.randn(B, L, hd, device=dev).to(BF); v = torch.randn(B, L, hd, device=dev).to(BF)
sc = torch.empty(B, L, L, device=dev, dtype=BF); o = torch.empty(B, L, hd, device=dev, dtype=BF)
timeit(lambda: ops.gemm(q, kk, sc, L, L, hd, B, 1, (L * hd, 0, hd, 1), (L * hd, 0, 1, hd), (L * L, 0, L)), 2.0 * B * L * L * hd, "QK^T 32x4096x4096x256")
timeit(lambda: ops.gemm(sc, v, o, L, hd, L, B, 1, (L * L, 0, L, 1), (L * hd, 0, hd, 1), (L * hd, 0, hd)), 2.0 * B * L * L * hd, "PV   32x4096x256x4096")
x = torch.randn(1, 1, 65536, 512, device=dev).to(BF)
w = torch.randn(512, 512, 1, 1, device=dev) / 512 ** 0.5
wp = ops.pack_weight(w, BF)
timeit(lambda: ops.conv_fprop(x, wp, None, None, 512, 1, 1), 2.0 * 65536 * 512 * 512, "Linear 65536x512x512")
w3 = torch.randn(1536, 512, 1, 1, device=dev) / 512 ** 0.5
wp3 = ops.pack_weight(w3, BF)
timeit(lambda: ops.conv_fprop(x, wp3, None, None, 1536, 1, 1), 2.0 * 65536 * 512 * 1536, "Linear 65536x1536x512")
