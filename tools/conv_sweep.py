"""Times every distinct dense shape of STC-UNet (N=16, 512x512) on the tcgen05 engine: fprop, dgrad, wgrad."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stc_unet_b200 as S
from stc_unet_b200 import ops
BF = torch.bfloat16
dev = torch.device("cuda:0")
N = int(os.environ.get("SWEEP_N", "16"))
# (name, Cin, Cout, HW, k)
SHAPES = [("L1 64->64 k3", 64, 64, 512, 3), ("L1 64->64 k5", 64, 64, 512, 5), ("L1 64->64 k7", 64, 64, 512, 7),
          ("L1 128->64 k3 (up4.0)", 128, 64, 512, 3),
          ("L2 64->128 k3", 64, 128, 256, 3), ("L2 128->128 k3", 128, 128, 256, 3), ("L2 128->128 k5", 128, 128, 256, 5),
          ("L2 128->128 k7", 128, 128, 256, 7), ("L2 256->64 k3 (up3.0)", 256, 64, 256, 3),
          ("L3 128->256 k3", 128, 256, 128, 3), ("L3 256->256 k3", 256, 256, 128, 3), ("L3 256->256 k5", 256, 256, 128, 5),
          ("L3 256->256 k7", 256, 256, 128, 7), ("L3 512->128 k3 (up2.0)", 512, 128, 128, 3),
          ("L4 256->512 k3", 256, 512, 64, 3), ("L4 512->512 k3", 512, 512, 64, 3), ("L4 1024->256 k3 (up1.0)", 1024, 256, 64, 3),
          ("L5 512->512 k3", 512, 512, 32, 3)]
only = os.environ.get("SWEEP_ONLY")
def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
rows = []
for name, ci, co, hw, k in SHAPES:
    if only and only not in name: continue
    x = torch.randn(N, hw, hw, ci, device=dev).to(BF)
    dy = torch.randn(N, hw, hw, co, device=dev).to(BF)
    w = torch.randn(co, ci, k, k, device=dev) / (ci * k * k) ** 0.5
    wp = ops.pack_weight(w, BF); wpt = ops.pack_weight(w, BF, transpose_flip=True)
    fl = 2.0 * N * hw * hw * ci * co * k * k
    t_f = timeit(lambda: ops.conv_fprop(x, wp, None, None, co, k, k))
    t_d = timeit(lambda: ops.conv_fprop(dy, wpt, None, None, ci, k, k))
    ws = torch.zeros(k * k * ci * co, device=dev)
    t_w = timeit(lambda: S._lib.lib.call("stc_conv_wgrad", x, dy, ws, N, hw, hw, ci, co, k, k, 1, 0, S._lib.stream_ptr()))
    rows.append((name, fl / 1e9, t_f, fl / t_f / 1e9, t_d, fl / t_d / 1e9, t_w, fl / t_w / 1e9))
    print("%-26s %8.1f GF | fprop %7.3f ms %6.0f TF/s | dgrad %7.3f ms %6.0f TF/s | wgrad %7.3f ms %6.0f TF/s" % rows[-1], flush=True)
    del x, dy
tf = sum(r[2] for r in rows); td = sum(r[4] for r in rows); tw = sum(r[6] for r in rows)
print("sum fprop %.2f ms dgrad %.2f ms wgrad %.2f ms" % (tf, td, tw))
