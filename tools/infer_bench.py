"""BASELINE.json configs[3] measured: STC-UNet sliding-window inference (test_cfg mode='slide', crop 256, stride 170: 9 windows per
512x512 slice) over a synthetic 512x512x64 volume with the integer confusion-matrix metrics, bf16, one B200.  The volume's 8-bit
slices are resident in HBM (and, for the e2e figure, copied from pinned host memory inside the timed region); the confusion matrix is
accumulated on the device and read back once per volume.  Prints one JSON line.  Also: --mode whole."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import stc_unet_b200 as S
from stc_unet_b200.metrics import ConfusionMeter

ap = argparse.ArgumentParser()
ap.add_argument("--slices", type=int, default=64)
ap.add_argument("--batch", type=int, default=8, help="slices per forward (x9 windows in slide mode)")
ap.add_argument("--mode", default="slide", choices=["slide", "whole"])
ap.add_argument("--volumes", type=int, default=3)
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
bcfg, hcfg = bench.model_cfg("stc", 3, "bf16")
tcfg = dict(mode="slide", crop_size=(256, 256), stride=(170, 170), max_windows_per_forward=72) if args.mode == "slide" else dict(mode="whole")
seg = S.EncoderDecoder(bcfg, hcfg, test_cfg=tcfg).to(dev)
seg.backbone.init_weights(); seg.decode_head.init_weights()
seg.eval()
S.ops.config.cache_eval_weights = True      # deployed checkpoint: weights are frozen, keep the folded / packed operands between forwards
seg.backbone.img_norm_cfg = dict(mean=[0.0], std=[255.0], to_rgb=False)     # uint8 HWC slices, normalised on the device
g = torch.Generator().manual_seed(5)
h_vol = torch.randint(0, 256, (args.slices, 512, 512, 3), dtype=torch.uint8, generator=g).pin_memory()
h_lab = torch.randint(0, 3, (args.slices, 512, 512), dtype=torch.uint8, generator=g).pin_memory()
d_vol, d_lab = h_vol.to(dev), h_lab.to(dev)


def run_volume(vol, lab, from_host):
    meter = ConfusionMeter(3, 255, device=dev)
    for s0 in range(0, args.slices, args.batch):
        img, lb = vol[s0:s0 + args.batch], lab[s0:s0 + args.batch]
        if from_host:
            img, lb = img.to(dev, non_blocking=True), lb.to(dev, non_blocking=True)
        pred = seg.inference_device(img)
        meter.update(pred, lb)
    return meter


def timed(from_host):
    vol, lab = (h_vol, h_lab) if from_host else (d_vol, d_lab)
    run_volume(vol, lab, from_host); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.volumes):
        m = run_volume(vol, lab, from_host)
        res = m.compute(["mIoU", "mDice"])        # D2H of the 3x3 int64 matrix: the volume's result
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.volumes, res, m


ms_dev, res, meter = timed(False)
ms_e2e, _, _ = timed(True)
wins = 9 if args.mode == "slide" else 1
gmac = 114.8 * 9 if args.mode == "slide" else 513.86        # forward GMAC per 256^2 crop / per 512^2 image
print(json.dumps(dict(metric=f"STC-UNet {args.mode} inference, 512x512x{args.slices} volume, bf16, confusion-matrix mIoU/Dice", unit="slices/s",
                      value=args.slices / ms_dev * 1e3, ms_per_volume=ms_dev, windows_per_slice=wins,
                      model_tflops_per_s=args.slices * gmac * 2e9 / (ms_dev * 1e-3) / 1e12,
                      e2e=dict(value=args.slices / ms_e2e * 1e3, ms_per_volume=ms_e2e, h2d_bytes_per_volume=h_vol.numel() + h_lab.numel(),
                               d2h_bytes_per_volume=72),
                      pixels_counted=int(meter.cm.sum()), mIoU=res["mIoU"], mDice=res["mDice"], n_gpus=1, data="synthetic")))
