"""Generates tests/golden/*.pt from the REFERENCE's own modules (imported from /root/reference through
oracle/ref_shim.py).  Run in the build container only:  python -m oracle.make_golden

Each fixture stores the seeded recipe (model kind, classes, seed, input size) plus the reference outputs: logits, the
three loss values, a few whole gradients and the L2 norm of every gradient, BN running statistics after one step, and
the area histograms of intersect_and_union on the argmax prediction.  Weights are NOT stored (STC-UNet has 41 M): the
test rebuilds them with torch.manual_seed(seed) + default init, which reproduces the reference's initialisation exactly
(checked here against the reference's state_dict hash)."""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sd_hash(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def make(kind: str, num_classes: int, size: int, batch: int, seed: int = 0):
    ns = ref_shim.load_reference()
    bb, hd = ref_shim.build_reference_model(kind == "stc", num_classes, dropout_ratio=0.0, seed=seed)
    init_hash = sd_hash({**{"b." + k: v for k, v in bb.state_dict().items()}, **{"h." + k: v for k, v in hd.state_dict().items()}})
    g = torch.Generator().manual_seed(seed + 100)
    img = torch.rand(batch, 3, size, size, generator=g)
    gt = torch.randint(0, num_classes, (batch, 1, size, size), generator=g)
    gt[:, :, :2] = 255
    bb.train(); hd.train()
    feats = bb(img)
    logits = hd(feats)
    losses = hd.losses(logits, gt)
    (losses["loss_bce"] + losses["loss_dice"]).backward()
    grads = {("backbone." + n): p.grad for n, p in bb.named_parameters()}
    grads.update({("decode_head." + n): p.grad for n, p in hd.named_parameters()})
    keep = ["decode_head.conv_seg.weight", "decode_head.conv_seg.bias", "decode_head.up4.conv.conv.4.weight", "backbone.inc.conv.conv.1.weight",
            "backbone.down4.down_conv.1.conv.4.bias"]
    pred = logits.argmax(dim=1).numpy()
    label = gt.squeeze(1).numpy().astype(np.uint8)
    areas = [torch.stack([t.double() for t in ns.intersect_and_union(pred[i], label[i], num_classes, 255)]) for i in range(batch)]
    fix = dict(kind=kind, num_classes=num_classes, size=size, batch=batch, seed=seed, init_hash=init_hash,
               logits=logits.detach().clone(), losses={k: float(v) for k, v in losses.items()},
               grads={k: grads[k].clone() for k in keep if k in grads},
               grad_norms={k: float(v.norm()) for k, v in grads.items()},
               running={("backbone." + k): v.clone() for k, v in bb.state_dict().items() if "running" in k and ("inc." in k or "down4" in k)},
               areas=torch.stack(areas).long(), state_keys={**{"backbone." + k: tuple(v.shape) for k, v in bb.state_dict().items()},
                                                             **{"decode_head." + k: tuple(v.shape) for k, v in hd.state_dict().items()}},
               torch_version=torch.__version__)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, f"{kind}_c{num_classes}_{size}.pt")
    torch.save(fix, path)
    print(path, os.path.getsize(path) // 1024, "KiB", {k: round(v, 6) for k, v in fix["losses"].items()})


if __name__ == "__main__":
    torch.set_num_threads(8)
    make("unet", 3, 32, 2)
    make("stc", 3, 32, 2)
    make("stc", 2, 48, 1)
