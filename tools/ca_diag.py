"""Diagnostic (GPU): CoordAtt inside Up (se=True) on offset-dominated inputs (the flip-free BN regime): error of every gradient
vs the fp64 oracle, for ours fp32 (fused upcat path and plain path) and for the oracle evaluated in fp32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import stc_unet_b200 as S
from stc_unet_b200 import ops
from oracle import stc_oracle as O
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
rel = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))
nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()
nchw = lambda t: t.permute(0, 3, 1, 2)

for (N, Cs, Cu, h, w) in ((2, 64, 64, 32, 32), (2, 256, 256, 8, 8)):
    torch.manual_seed(0)
    H, W = 2 * h, 2 * w
    C = Cs + Cu
    ca = S.CoordAtt(C, C).to(dev)
    skip = 4 + 0.25 * torch.randn(N, Cs, H, W, device=dev)
    low = 4 + 0.25 * torch.randn(N, Cu, h, w, device=dev)
    go = torch.randn(N, C, H, W, device=dev)

    def run_oracle(dt):
        sd = {"ca." + k: (v.detach().to(dt) if v.is_floating_point() else v.detach().clone()).requires_grad_(v.is_floating_point() and "running" not in k)
              for k, v in ca.state_dict().items()}
        s, l = skip.to(dt).requires_grad_(True), low.to(dt).requires_grad_(True)
        x = torch.cat([s, F.interpolate(l, scale_factor=2, mode="bilinear", align_corners=True)], 1)
        out = O.coord_att(sd, "ca", x, True, None) + x
        out.backward(go.to(dt))
        return dict(out=out.detach(), ds=s.grad, dl=l.grad, **{k[3:]: v.grad for k, v in sd.items() if v.requires_grad})

    def run_ours(fused):
        ca.zero_grad()
        s, l = nhwc(skip).requires_grad_(True), nhwc(low).requires_grad_(True)
        if fused:
            out = ops.upcat_coordatt(s, l, True, ca.attention)
        else:
            out = ca.forward_add(ops.upcat(s, l, True))
        gs, gl = torch.autograd.grad(out, (s, l), nhwc(go), retain_graph=True)
        out.backward(nhwc(go))
        return dict(out=nchw(out.detach()), ds=nchw(gs), dl=nchw(gl), **{k: p.grad.clone() for k, p in ca.named_parameters()})

    r64, r32 = run_oracle(torch.float64), run_oracle(torch.float32)
    of, op = run_ours(True), run_ours(False)
    print(f"=== N={N} Cs={Cs} Cu={Cu} {H}x{W}")
    for k in r64:
        if k == "conv1.bias":
            continue
        print("  %-16s torch32 %.2e  ours-fused %.2e  ours-plain %.2e" % (k, rel(r32[k], r64[k]), rel(of[k], r64[k]), rel(op[k], r64[k])))
