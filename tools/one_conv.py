"""Runs one conv shape a few times (for ncu).  usage: one_conv.py Cin Cout HW k [N] [mode: fprop|wgrad|both]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stc_unet_b200 as S
from stc_unet_b200 import ops
ci, co, hw, k = (int(a) for a in sys.argv[1:5])
N = int(sys.argv[5]) if len(sys.argv) > 5 else 16
mode = sys.argv[6] if len(sys.argv) > 6 else "fprop"
BF = torch.bfloat16; dev = torch.device("cuda:0")
x = torch.randn(N, hw, hw, ci, device=dev).to(BF)
dy = torch.randn(N, hw, hw, co, device=dev).to(BF)
w = torch.randn(co, ci, k, k, device=dev) / (ci * k * k) ** 0.5
wp = ops.pack_weight(w, BF)
ws = torch.zeros(k * k * ci * co, device=dev)
for _ in range(4):
    if mode in ("fprop", "both"):
        y = ops.conv_fprop(x, wp, None, None, co, k, k)
    if mode in ("wgrad", "both"):
        S._lib.lib.call("stc_conv_wgrad", x, dy, ws, N, hw, hw, ci, co, k, k, 1, 0, S._lib.stream_ptr())
torch.cuda.synchronize()
print("ok")
