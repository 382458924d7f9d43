"""Where a halo-conv CTA spends its time: per-CTA clock counters written by umma_convh_kernel when stc_debug_profile() is armed.
   usage: convh_prof.py [Cin Cout HW k] ...   (default: the 3x3 shapes of the STC-UNet step)
   counters (clocks, per CTA): 0 MMA-issuer loop total | 1 wait accumulator free | 2 wait input segment | 3 wait weight tile | 4 strips
                               5 epilogue wait accumulator full | 6 epilogue work | 8 A producer wait slot | 9 B producer wait slot"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stc_unet_b200 as S
from stc_unet_b200 import ops
BF = torch.bfloat16; dev = torch.device("cuda:0")
shapes = [(64, 64, 512, 3), (128, 128, 256, 3), (256, 256, 128, 3), (64, 64, 512, 7), (128, 128, 256, 7), (128, 64, 512, 3)]
a = [int(v) for v in sys.argv[1:]]
if a: shapes = [tuple(a[i:i + 4]) for i in range(0, len(a), 4)]
prof = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
for (ci, co, hw, k) in shapes:
    x = torch.randn(16, hw, hw, ci, device=dev).to(BF)
    w = torch.randn(co, ci, k, k, device=dev) / (ci * k * k) ** 0.5
    wp = ops.pack_weight(w, BF)
    for _ in range(3): ops.conv_fprop(x, wp, None, None, co, k, k)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): ops.conv_fprop(x, wp, None, None, co, k, k)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    prof.zero_()
    S._lib.lib.call("stc_debug_profile", prof)
    ops.conv_fprop(x, wp, None, None, co, k, k)
    torch.cuda.synchronize()
    S._lib.lib.call("stc_debug_profile", None)
    c = prof.view(148, 16).double()
    m = c.mean(0)
    lead = c[c[:, 0] > 0]          # cta_group::2: only the pair leaders issue MMAs
    m[:5] = lead.mean(0)[:5]
    fl = 2.0 * 16 * hw * hw * ci * co * k * k
    print(f"{ci}->{co} k{k} @{hw}: {ms:.3f} ms {fl / ms / 1e9:.0f} TF/s | per CTA (mean clocks): loop {m[0]:.0f} strips {m[4]:.1f} | MMA waits: acc-free {m[1]:.0f} ({100 * m[1] / m[0]:.0f}%) segment {m[2]:.0f} ({100 * m[2] / m[0]:.0f}%) weights {m[3]:.0f} ({100 * m[3] / m[0]:.0f}%) | epilogue: wait-full {m[5]:.0f} work {m[6]:.0f} ({100 * m[6] / m[0]:.0f}%) | producers: A wait-slot {m[8]:.0f} B wait-slot {m[9]:.0f} | loop min/max {lead[:, 0].min():.0f}/{lead[:, 0].max():.0f}", flush=True)

# the short-K GEMMs of the attention blocks through the plain tcgen05 GEMM (same counter layout; 2 = wait smem stage)
def gemm_prof(name, fn, flops):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): fn()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    prof.zero_()
    S._lib.lib.call("stc_debug_profile", prof)
    fn(); torch.cuda.synchronize()
    S._lib.lib.call("stc_debug_profile", None)
    c = prof.view(148, 16).double(); c = c[c[:, 0] > 0]; m = c.mean(0)
    print(f"{name}: {ms:.3f} ms {flops / ms / 1e9:.0f} TF/s | per MMA-issuing CTA: loop {m[0]:.0f} tiles {m[4]:.1f} ({m[0] / m[4]:.0f} clk/tile) | MMA waits: acc-free {m[1]:.0f} ({100 * m[1] / m[0]:.0f}%) smem stage {m[2]:.0f} ({100 * m[2] / m[0]:.0f}%) | epilogue: wait-full {m[5]:.0f} work {m[6]:.0f} ({100 * m[6] / m[0]:.0f}%, {m[6] / m[4]:.0f} clk/tile)", flush=True)
if not a:
    L, hd, B = 4096, 256, 32
    q = torch.randn(B, L, hd, device=dev).to(BF); kk = torch.randn(B, L, hd, device=dev).to(BF); v = torch.randn(B, L, hd, device=dev).to(BF)
    sc = torch.empty(B, L, L, device=dev, dtype=BF); o = torch.empty(B, L, hd, device=dev, dtype=BF)
    gemm_prof("QK^T 32x4096x4096x256", lambda: ops.gemm(q, kk, sc, L, L, hd, B, 1, (L * hd, 0, hd, 1), (L * hd, 0, 1, hd), (L * L, 0, L)), 2.0 * B * L * L * hd)
    gemm_prof("PV   32x4096x256x4096", lambda: ops.gemm(sc, v, o, L, hd, L, B, 1, (L * L, 0, L, 1), (L * hd, 0, hd, 1), (L * hd, 0, hd)), 2.0 * B * L * L * hd)
    x = torch.randn(1, 1, 65536, 512, device=dev).to(BF)
    w = torch.randn(512, 512, 1, 1, device=dev) / 512 ** 0.5
    wp = ops.pack_weight(w, BF)
    gemm_prof("Linear 65536x512x512", lambda: ops.conv_fprop(x, wp, None, None, 512, 1, 1), 2.0 * 65536 * 512 * 512)
