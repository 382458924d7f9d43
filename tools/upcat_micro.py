"""Micro-run of the fused Up.forward pieces at one level's shape (default: up4, N=16, 64+64 channels at 512x512) for ncu."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stc_unet_b200 as S  # noqa
from stc_unet_b200 import ops

N, C, H = int(os.environ.get("N", 16)), int(os.environ.get("C", 64)), int(os.environ.get("H", 512))
torch.manual_seed(0)
skip = torch.randn(N, H, H, C, device="cuda", dtype=torch.bfloat16, requires_grad=True)
low = torch.randn(N, H // 2, H // 2, C, device="cuda", dtype=torch.bfloat16, requires_grad=True)
att = lambda y, n, h, w: torch.sigmoid(y)
for it in range(2):
    out = ops.upcat_coordatt(skip, low, True, att)
    out.backward(torch.ones_like(out))
    skip.grad = low.grad = None
torch.cuda.synchronize()
print("ok")
