"""Training-step time of the OTHER model families of BASELINE.json's configs on one B200 (bf16, batch 16 of 3x512x512, Trainer =
fwd + loss + bwd + fused Adam, whole-step CUDA graph): plain U-Net (my_config/U-Net.py), mmseg UNet-S5-D16 + FCNHead (family B) and
UNet++ (my_config/UNet++.py, config 5).  One JSON line per model."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import stc_unet_b200 as S
from stc_unet_b200.train import Trainer

dev = torch.device("cuda", 0)
LOSS = bench.LOSS_CFG
norm = dict(type="BN", requires_grad=True)


def make(kind):
    torch.manual_seed(0)
    if kind == "unet":
        b, h = bench.model_cfg("unet", 3, "bf16")
        return S.EncoderDecoder(b, h), 123.70
    if kind == "unet_b":
        b = dict(type="UNet", in_channels=3, base_channels=64, num_stages=5, strides=(1,) * 5, enc_num_convs=(2,) * 5, dec_num_convs=(2,) * 4,
                 downsamples=(True,) * 4, enc_dilations=(1,) * 5, dec_dilations=(1,) * 4, with_cp=False, conv_cfg=None, norm_cfg=norm,
                 act_cfg=dict(type="ReLU"), upsample_cfg=dict(type="InterpConv"), norm_eval=False, compute_dtype="bf16")
        h = dict(type="FCNHead", in_channels=64, in_index=4, channels=64, num_convs=1, concat_input=False, dropout_ratio=0.1, num_classes=3,
                 norm_cfg=norm, align_corners=False, loss_decode=dict(type="CrossEntropyLoss", use_sigmoid=False, loss_weight=1.0))
        return S.EncoderDecoder(b, h), 207.0 - 14.5 / 2       # encoder 68.1 + decoder 124.6 + main head (aux head not built here)
    seg = S.build_segmentor(dict(type="EncoderDecoderFull", decode_head=dict(type="UnetPlusPlus", num_classes=3, norm_cfg=norm, loss_decode=LOSS,
                                                                              dropout_ratio=0.1, compute_dtype="bf16")))
    return seg, 376.0


for kind in (sys.argv[1:] or ["unet", "unet_b", "unetpp"]):
    seg, gmac = make(kind)
    seg = seg.to(dev)
    for m in seg.modules():
        if hasattr(m, "init_weights") and m is not seg:
            pass
    seg.train()
    tr = Trainer(seg, lr=1e-5)
    g = torch.Generator().manual_seed(1)
    img = torch.rand(16, 3, 512, 512, generator=g).to(dev)
    gt = torch.randint(0, 3, (16, 1, 512, 512), generator=g).to(dev)
    tr.capture(img, gt)
    for _ in range(3):
        lv = tr.step_graph(img, gt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 10
    for _ in range(K):
        lv = tr.step_graph(img, gt)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(json.dumps(dict(model=kind, ms_per_step=ms, img_per_s=16 / ms * 1e3, loss=float(lv["loss"]),
                          effective_tflops=16 * gmac * 6e9 / (ms * 1e-3) / 1e12, fwd_gmac_per_img=gmac, batch=16, dtype="bf16", cuda_graph=True)), flush=True)
    del tr, seg
    torch.cuda.empty_cache()
