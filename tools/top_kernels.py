"""One launch of each dominant tcgen05 kernel at its heaviest in-step shape (for `ncu --set full`): halo fprop L1 64->64 k7 and k3,
halo wgrad L1 64->64 k7, Linear 512x512 @65536 rows (fprop + wgrad), level-2 128->128 (N = 128: no shared-memory bound), conv + BN
statistics, QK^T (K = 256: epilogue / TMEM read-out bound) and PV (K = 4096: the cta_group::2 pair kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stc_unet_b200 as S
from stc_unet_b200 import ops
BF = torch.bfloat16; dev = torch.device("cuda:0")
def conv(ci, co, hw, k, N=16, wgrad=False):
    x = torch.randn(N, hw, hw, ci, device=dev).to(BF); dy = torch.randn(N, hw, hw, co, device=dev).to(BF)
    w = torch.randn(co, ci, k, k, device=dev) / (ci * k * k) ** 0.5
    wp = ops.pack_weight(w, BF); ws = torch.zeros(k * k * ci * co, device=dev)
    for _ in range(2):
        if wgrad: S._lib.lib.call("stc_conv_wgrad", x, dy, ws, N, hw, hw, ci, co, k, k, 1, 0, S._lib.stream_ptr())
        else: ops.conv_fprop(x, wp, None, None, co, k, k)
    torch.cuda.synchronize()
conv(64, 64, 512, 7); conv(64, 64, 512, 3); conv(64, 64, 512, 7, wgrad=True); conv(512, 512, 64, 1); conv(512, 512, 64, 1, wgrad=True)
conv(128, 128, 256, 7); conv(128, 128, 256, 3, wgrad=True)      # level 2: N = 128 (the shared-memory bound of N = 64 is gone)
def conv_stats(ci, co, hw, k, N=16):                                # conv + BN statistics out of the epilogue (stc_conv_fprop_bnstats)
    x = torch.randn(N, hw, hw, ci, device=dev).to(BF)
    w = torch.randn(co, ci, k, k, device=dev) / (ci * k * k) ** 0.5
    wp = ops.pack_weight(w, BF)
    for _ in range(2):
        ops.conv_fprop_bnstats(x, wp, None, co, k, k)
    torch.cuda.synchronize()
conv_stats(64, 64, 512, 7)
L, hd, B = 4096, 256, 32
q = torch.randn(B, L, hd, device=dev).to(BF); kk = torch.randn(B, L, hd, device=dev).to(BF); v = torch.randn(B, L, hd, device=dev).to(BF)
sc = torch.empty(B, L, L, device=dev, dtype=BF); o = torch.empty(B, L, hd, device=dev, dtype=BF)
for _ in range(2):
    ops.gemm(q, kk, sc, L, L, hd, B, 1, (L * hd, 0, hd, 1), (L * hd, 0, 1, hd), (L * L, 0, L))
for _ in range(2):
    ops.gemm(sc, v, o, L, hd, L, B, 1, (L * L, 0, L, 1), (L * hd, 0, hd, 1), (L * hd, 0, hd))
conv(512, 512, 64, 3, wgrad=True)                                   # per-tap wgrad (W < 128), division-free producer
# CTA-pair (cta_group::2) halo kernels, further shapes: 5x5 at level 1, BN = 256 at level 3, ci-chunk-pair wgrad at level 2, 3x3 wgrad at level 1
conv(64, 64, 512, 5); conv(256, 256, 128, 3); conv(128, 128, 256, 7, wgrad=True); conv(64, 64, 512, 3, wgrad=True)
torch.cuda.synchronize(); print("ok")
