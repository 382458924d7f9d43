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
if os.environ.get("PROBE_CUBLAS", "0") == "1":
    # library yardstick (torch.matmul -> cuBLASLt) on the same shapes, plus cuDNN on the conv shapes that bound the step: measurement only
    timeit(lambda: torch.bmm(q, kk.transpose(1, 2), out=sc), 2.0 * B * L * L * hd, "cuBLAS QK^T 32x4096x4096x256")
    timeit(lambda: torch.bmm(sc, v, out=o), 2.0 * B * L * L * hd, "cuBLAS PV   32x4096x256x4096")
    x2 = x.view(65536, 512); w2 = torch.randn(512, 512, device=dev).to(BF); y2 = torch.empty(65536, 512, device=dev, dtype=BF)
    timeit(lambda: torch.mm(x2, w2, out=y2), 2.0 * 65536 * 512 * 512, "cuBLAS Linear 65536x512x512")
    w23 = torch.randn(512, 1536, device=dev).to(BF); y23 = torch.empty(65536, 1536, device=dev, dtype=BF)
    timeit(lambda: torch.mm(x2, w23, out=y23), 2.0 * 65536 * 512 * 1536, "cuBLAS Linear 65536x1536x512")
    for (M, N, K) in ((4194304, 64, 576), (4194304, 64, 3136), (1048576, 128, 1152), (1048576, 128, 6272), (4194304, 64, 64), (4194304, 128, 128)):
        a = torch.randn(M, K, device=dev).to(BF); b = torch.randn(K, N, device=dev).to(BF); c = torch.empty(M, N, device=dev, dtype=BF)
        timeit(lambda: torch.mm(a, b, out=c), 2.0 * M * N * K, f"cuBLAS GEMM {M}x{N}x{K}")
        del a, b, c
    import torch.nn.functional as F
    torch.backends.cudnn.benchmark = True
    for (C, Co, k, HW) in ((64, 64, 3, 512), (64, 64, 7, 512), (128, 128, 3, 256), (128, 128, 7, 256), (256, 256, 3, 128)):
        xi = torch.randn(16, C, HW, HW, device=dev).to(BF).contiguous(memory_format=torch.channels_last)
        wi = torch.randn(Co, C, k, k, device=dev).to(BF).contiguous(memory_format=torch.channels_last)
        fl = 2.0 * 16 * HW * HW * C * Co * k * k
        timeit(lambda: F.conv2d(xi, wi, padding=k // 2), fl, f"cuDNN fprop {C}->{Co} k{k} @{HW}")
        yi = F.conv2d(xi, wi, padding=k // 2)
        timeit(lambda: torch.ops.aten.convolution_backward(yi, xi, wi, None, (1, 1), (k // 2, k // 2), (1, 1), False, (0, 0), 1, (False, True, False)), fl, f"cuDNN wgrad {C}->{Co} k{k} @{HW}")
        timeit(lambda: torch.ops.aten.convolution_backward(yi, xi, wi, None, (1, 1), (k // 2, k // 2), (1, 1), False, (0, 0), 1, (True, False, False)), fl, f"cuDNN dgrad {C}->{Co} k{k} @{HW}")
        del xi, wi, yi
