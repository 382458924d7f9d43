"""One eager STC-UNet training step (bf16, N=16, 512x512) bracketed by cudaProfilerStart/Stop, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_raw.csv python tools/ncu_step.py
The three steps before it (StepCache record / finalize / replay) run unprofiled.  tools/ncu_launch_list.py turns the raw CSV into
profiles/<round>_launches_*.csv + a per-kernel summary."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import stc_unet_b200 as S  # noqa: E402
from stc_unet_b200.train import Trainer  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
bcfg, hcfg = bench.model_cfg("stc", 3, "bf16")
seg = S.EncoderDecoder(bcfg, hcfg).to(dev)
seg.backbone.init_weights(); seg.decode_head.init_weights()
seg.train()
trainer = Trainer(seg, lr=1e-5, betas=(0.9, 0.999))
img = torch.rand(16, 3, 512, 512, device=dev)
gt = torch.randint(0, 3, (16, 1, 512, 512), device=dev)
for _ in range(3):
    trainer.step(img, gt)
torch.cuda.synchronize()
torch.cuda.profiler.start()
lv = trainer.step(img, gt)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(lv["loss"]))
