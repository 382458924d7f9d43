"""Diagnostic (GPU): 1x1 conv + train-mode BN + h_swish on offset-dominated inputs; gradient errors vs fp64."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import stc_unet_b200 as S
from stc_unet_b200 import ops
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
rel = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))
torch.manual_seed(0)
for (N, R, C, M, off) in ((2, 128, 128, 32, 4.0), (2, 128, 128, 32, 0.0), (2, 32, 512, 128, 4.0)):
    conv = torch.nn.Conv2d(C, M, 1).to(dev)
    bn = torch.nn.BatchNorm2d(M).to(dev)
    y = (off + 0.03 * torch.randn(N, R, 1, C, device=dev))          # NHWC
    go = torch.randn(N, R, 1, M, device=dev)

    def ref(dt):
        w, b, g, be = (t.detach().to(dt).requires_grad_(True) for t in (conv.weight, conv.bias, bn.weight, bn.bias))
        x = y.permute(0, 3, 1, 2).to(dt).requires_grad_(True)
        z = F.conv2d(x, w, b)
        o = F.batch_norm(z, None, None, g, be, True, 0.1, 1e-5)
        o = o * torch.clamp(o + 3, 0, 6) / 6
        o.backward(go.permute(0, 3, 1, 2).to(dt))
        return dict(out=o.detach().permute(0, 2, 3, 1), dx=x.grad.permute(0, 2, 3, 1), dw=w.grad, dg=g.grad, db=be.grad)

    def ours():
        conv.zero_grad(); bn.zero_grad()
        x = y.clone().requires_grad_(True)
        o = ops.conv_bn_act(x, conv, bn, S.modules.ACT_HSWISH, True)
        o.backward(go)
        return dict(out=o.detach(), dx=x.grad, dw=conv.weight.grad.clone(), dg=bn.weight.grad.clone(), db=bn.bias.grad.clone())
    r64, r32, o32 = ref(torch.float64), ref(torch.float32), ours()
    print(f"=== rows {N*R} C {C} -> {M} offset {off}")
    for k in r64:
        print("  %-4s torch32 %.2e ours %.2e" % (k, rel(r32[k], r64[k]), rel(o32[k], r64[k])))
