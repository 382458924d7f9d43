"""Diagnostic (GPU): error of every gradient vs an fp64 oracle, for ours (fp32/bf16) and for torch fp32 / autocast-bf16."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import stc_oracle as O
from tests.test_model_gpu import build
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False

def rel(a, b): return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))

_build = build
def build(stc, C, dt):
    bb, hd = _build(stc, C, dt)
    if os.environ.get("DIAG_POSBN"):
        for m in list(bb.modules()) + list(hd.modules()):
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                m.weight.data.fill_(0.25); m.bias.data.fill_(4.0)
    return bb, hd

def oracle(bb, hd, img, gt, dt, autocast=False):
    conv = lambda v: v.detach().clone().to(dt) if v.is_floating_point() else v.detach().clone()
    bsd = {k: conv(v).requires_grad_(v.is_floating_point() and "running" not in k) for k, v in bb.state_dict().items()}
    hsd = {k: conv(v).requires_grad_(v.is_floating_point() and "running" not in k) for k, v in hd.state_dict().items()}
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        feats = O.backbone_forward(bsd, img.to(dt), True, None)
        logits = O.head_forward(hsd, feats, True, None)
    out = O.losses(logits.float() if autocast else logits, gt)
    (out["loss_bce"] + out["loss_dice"]).backward()
    grads = {("b", k): v.grad for k, v in bsd.items() if v.requires_grad} | {("h", k): v.grad for k, v in hsd.items() if v.requires_grad}
    return logits.detach(), grads, out

size = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for stc in (False, True):
    g = torch.Generator().manual_seed(123)
    img = torch.rand(2, 3, size, size, generator=g).cuda(); gt = torch.randint(0, 3, (2, 1, size, size), generator=g).cuda()
    if os.environ.get("DIAG_STRUCT"):
        gt = (img.mean(1, keepdim=True) * 3).floor().clamp(0, 2).long()
    bb, hd = build(stc, 3, "fp32")
    l64, g64, o64 = oracle(bb, hd, img, gt, torch.float64)
    l32, g32, _ = oracle(bb, hd, img, gt, torch.float32)
    lac, gac, _ = oracle(bb, hd, img, gt, torch.float32, autocast=True)
    res = {}
    for dt in ("fp32", "bf16"):
        b2, h2 = build(stc, 3, dt)
        feats = b2(img); losses = h2.forward_train(feats, None, gt, None)
        (losses["loss_bce"] + losses["loss_dice"]).backward()
        with torch.no_grad():
            b3, h3 = build(stc, 3, dt); lg = h3(b3(img))
        grads = {("b", k): p.grad for k, p in b2.named_parameters()} | {("h", k): p.grad for k, p in h2.named_parameters()}
        res[dt] = (lg, grads, losses)
    print(f"=== stc={stc} size={size}")
    print(" logits rel vs fp64: torch32 %.2e autocast %.2e ours32 %.2e oursbf16 %.2e" % (rel(l32, l64), rel(lac, l64), rel(res["fp32"][0], l64), rel(res["bf16"][0], l64)))
    print(" losses fp64", {k: float(v) for k, v in o64.items()}, "ours32", {k: float(v) for k, v in res["fp32"][2].items()}, "oursbf16", {k: float(v) for k, v in res["bf16"][2].items()})
    rows = []
    for k in g64:
        if k[1].endswith("bias") and ("conv.conv" in k[1] or ".convs." in k[1] or "ca.conv1" in k[1]):
            continue
        rows.append((rel(res["fp32"][1][k], g64[k]), rel(g32[k], g64[k]), rel(res["bf16"][1][k], g64[k]), rel(gac[k], g64[k]), k[1]))
    rows.sort(reverse=True)
    print(" worst by ours32: (ours32, torch32, oursbf16, autocast, name)")
    for r in rows[:12]: print("   %.2e %.2e %.2e %.2e %s" % r)
    rows.sort(key=lambda r: -r[2])
    print(" worst by oursbf16:")
    for r in rows[:12]: print("   %.2e %.2e %.2e %.2e %s" % r)
    if os.environ.get("DIAG_ALL"):
        order = {k[1]: i for i, k in enumerate(g64)}
        rows.sort(key=lambda r: order[r[4]])
        print(" all (forward order): ours32 torch32 oursbf16 autocast")
        for r in rows: print("   %.2e %.2e %.2e %.2e %s" % r)
    import statistics
    print(" median: ours32 %.2e torch32 %.2e oursbf16 %.2e autocast %.2e" % tuple(statistics.median(r[i] for r in rows) for i in range(4)))
