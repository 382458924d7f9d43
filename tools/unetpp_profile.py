import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import stc_unet_b200 as S
from stc_unet_b200 import ops
from stc_unet_b200.train import Trainer
dev = torch.device("cuda", 0)
norm = dict(type="BN", requires_grad=True)
seg = S.build_segmentor(dict(type="EncoderDecoderFull", decode_head=dict(type="UnetPlusPlus", num_classes=3, norm_cfg=norm, loss_decode=bench.LOSS_CFG,
                                                                          dropout_ratio=0.1, compute_dtype="bf16"))).to(dev).train()
tr = Trainer(seg, lr=1e-5)
img = torch.rand(16, 3, 512, 512, device=dev); gt = torch.randint(0, 3, (16, 1, 512, 512), device=dev)
for _ in range(3): tr.step(img, gt)
# wrap lib.call to record shapes of conv calls
recs = []
orig = S._lib.lib.call
def call(name, *a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = orig(name, *a); e1.record()
    key = name
    if name in ("stc_conv_fprop", "stc_conv_fprop_bnstats"):
        off = 5 if name == "stc_conv_fprop" else 4
        key = f"{name} HxW={a[off+1]}x{a[off+2]} Cin={a[off+3]} Cout={a[off+4]} k={a[off+5]} eng={S._lib.lib.raw('stc_dense_last_engine')()}"
    elif name == "stc_conv_wgrad":
        key = f"{name} HxW={a[4]}x{a[5]} Cin={a[6]} Cout={a[7]} k={a[8]} eng={S._lib.lib.raw('stc_dense_last_engine')()}"
    recs.append((key, e0, e1)); return r
S._lib.lib.call = call
tr.step(img, gt); torch.cuda.synchronize()
del S._lib.lib.__dict__["call"]
agg = collections.OrderedDict()
for k, a, b in recs:
    d = agg.setdefault(k, [0, 0.0]); d[0] += 1; d[1] += a.elapsed_time(b)
tot = sum(v[1] for v in agg.values()); print("total", tot)
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]: print("%8.2f ms %3d  %s" % (ms, n, k))
