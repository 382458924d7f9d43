#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: STC-UNet training throughput (img/s) at 512x512, bf16, batch 16 per B200.

  python bench.py --gpus N --steps K --warmup W          (N>1: launched per rank by torch.distributed.run)
  python bench.py --impl reference ...                   (the reference path's arithmetic on the host CPU cores)

A "step" = one full training iteration of my_config/STC-UNet.py's model on one synthetic batch:
forward (UnetBackbone + KSA + Transformer, UnetHead + CoordAtt, cls_seg with Dropout2d) + CE/Dice loss + backward
+ (N>1) SyncBN statistic exchange and bucketed gradient all-reduce + fused Adam step.  Nothing is skipped.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is obtained.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "STC-UNet train img/s @512x512 bf16 (fwd+loss+bwd+allreduce+Adam)"
# SURVEY.md §8(d), forward MACs per 512x512 image (unet_b: encoder 68.1 + decoder 124.6 + main FCN head; unetpp: derived, unverified)
FWD_GMAC_PER_IMG = {"stc": 513.86, "unet": 123.70, "unet_b": 199.75, "unetpp": 376.0}
MODEL_NAME = {"stc": "my_config/STC-UNet.py", "unet": "my_config/U-Net.py", "unet_b": "configs/_base_/models/fcn_unet_s5-d16.py (UNet-S5-D16 + FCNHead)",
              "unetpp": "my_config/UNet++.py (EncoderDecoderFull + UnetPlusPlus; parity unpinned)"}
SLIDE_GMAC_PER_SLICE = 114.8 * 9                      # 256x256 crop forward x 9 windows per 512x512 slice
LOSS_CFG = [dict(type="CrossEntropyLoss", use_sigmoid=False, loss_name="loss_bce", loss_weight=1.0),
            dict(type="DiceLoss", loss_name="loss_dice", loss_weight=1.0)]
# kernel launches per C-ABI call (default 1); memsets are not counted
LAUNCHES = {"stc_bn_reduce": 2, "stc_bn_bwd_reduce": 2, "stc_seg_loss_fwd": 2, "stc_upcat_bwd": 2}


def make_segmentor(kind: str, num_classes: int, dtype: str, **head_kw):
    """The segmentor of one of BASELINE.json's configs (random init; the caller moves it to the device)."""
    import stc_unet_b200 as S
    norm = dict(type="BN", requires_grad=True)
    if kind in ("stc", "unet"):
        bcfg, hcfg = model_cfg(kind, num_classes, dtype)
        seg = S.EncoderDecoder(bcfg, dict(hcfg, **head_kw))
        seg.backbone.init_weights(); seg.decode_head.init_weights()
        return seg
    if kind == "unet_b":
        b = dict(type="UNet", in_channels=3, base_channels=64, num_stages=5, strides=(1,) * 5, enc_num_convs=(2,) * 5, dec_num_convs=(2,) * 4,
                 downsamples=(True,) * 4, enc_dilations=(1,) * 5, dec_dilations=(1,) * 4, with_cp=False, conv_cfg=None, norm_cfg=norm,
                 act_cfg=dict(type="ReLU"), upsample_cfg=dict(type="InterpConv"), norm_eval=False, compute_dtype=dtype)
        h = dict(type="FCNHead", in_channels=64, in_index=4, channels=64, num_convs=1, concat_input=False, dropout_ratio=0.1,
                 num_classes=num_classes, norm_cfg=norm, align_corners=False,
                 loss_decode=dict(type="CrossEntropyLoss", use_sigmoid=False, loss_weight=1.0))
        seg = S.EncoderDecoder(b, dict(h, **head_kw))
        seg.backbone.init_weights(); seg.decode_head.init_weights()
        return seg
    seg = S.build_segmentor(dict(type="EncoderDecoderFull", decode_head=dict(type="UnetPlusPlus", num_classes=num_classes, norm_cfg=norm,
                                                                              loss_decode=LOSS_CFG, compute_dtype=dtype, **head_kw)))
    seg.decode_head.init_weights()
    return seg


def oracle_step_factory(kind, seg, img, gt, autocast_dtype=None, optimizer=False):
    """fwd + loss + bwd (+ torch Adam) of the ORACLE restatement on the tensors' device, from `seg`'s state_dict.  This is the reference
    path's arithmetic in plain PyTorch: on the CPU it is the cpu_baseline / --impl reference arm; on the GPU under autocast(bf16) it is
    the informative `gpu_eager_baseline` (cuDNN / cuBLAS eager: "just run the reference on a B200")."""
    from oracle import stc_oracle as O
    dev = img.device
    def leaves(sd):
        return {k: v.detach().to(dev).clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    if kind in ("stc", "unet", "unet_b"):
        bsd, hsd = leaves(seg.backbone.state_dict()), leaves(seg.decode_head.state_dict())
        params = [v for d in (bsd, hsd) for v in d.values() if v.requires_grad]
    else:
        hsd = leaves(seg.decode_head.state_dict())
        bsd = {}
        params = [v for v in hsd.values() if v.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-5, betas=(0.9, 0.999)) if optimizer else None

    def step():
        for v in params:
            v.grad = None
        with torch.autocast(dev.type, dtype=autocast_dtype, enabled=autocast_dtype is not None):
            if kind in ("stc", "unet"):
                logits = O.head_forward(hsd, O.backbone_forward(bsd, img, True, {}), True, {})
            elif kind == "unet_b":
                logits = O.fcn_head_forward(hsd, O.unet_b_forward(bsd, img, True, {}), 4, True, {})
            else:
                logits = O.unetpp_forward(hsd, img, True, {})
        out = O.losses(logits.float(), gt)
        loss = out["loss_bce"] if kind == "unet_b" else out["loss_bce"] + out["loss_dice"]
        loss.backward()
        if opt is not None:
            opt.step()
        return loss
    return step


def model_cfg(kind: str, num_classes: int, dtype: str):
    if kind == "stc":   # my_config/STC-UNet.py:2-19 (num_classes per BASELINE.json configs: 3)
        backbone = dict(type="UnetBackbone", in_channels=3, context_layer="kernelselect", transformer_block=True,
                        channel_list=[64, 128, 256, 512], compute_dtype=dtype)
        head = dict(type="UnetHead", se=True, num_classes=num_classes, channels=64, threshold=0.2,
                    norm_cfg=dict(type="BN", requires_grad=True), loss_decode=LOSS_CFG)
    else:               # my_config/U-Net.py:2-17
        backbone = dict(type="UnetBackbone", in_channels=3, channel_list=[64, 128, 256, 512], compute_dtype=dtype)
        head = dict(type="UnetHead", num_classes=num_classes, channels=64, threshold=0.2,
                    norm_cfg=dict(type="BN", requires_grad=True), loss_decode=LOSS_CFG)
    return backbone, head


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        if not rows:
            return None
        sm = [float(r[0]) for r in rows]
        reasons = [n for i, n in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap"))
                   if any(r[i].lower().startswith("active") for r in rows)]
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=float(rows[0][1]), power_w_max=max(float(r[2]) for r in rows),
                    samples=len(rows), reasons=reasons)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops_sustained=d.get("bf16_tflops_sustained"), tflops_burst=d.get("bf16_tflops"), hbm_gbs=d.get("hbm_gbs"), source="measured")
    return dict(tflops_sustained=1400.0, tflops_burst=1590.0, hbm_gbs=6650.0, source="fallback")


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference path's arithmetic (oracle port; /root/reference itself is absent on the GPU box)
# ------------------------------------------------------------------------------------------------
def cpu_step_factory(kind, num_classes, size, batch):
    torch.manual_seed(0)
    seg = make_segmentor(kind, num_classes, "fp32", dropout_ratio=0.0)
    g = torch.Generator().manual_seed(0)
    img = torch.rand(batch, 3, size, size, generator=g)
    gt = torch.randint(0, num_classes, (batch, 1, size, size), generator=g)
    step = oracle_step_factory(kind, seg, img, gt)
    return lambda: float(step())


def cpu_slide_factory(num_classes, slices):
    """Config 4 on the host cores: the oracle's slide_inference (9 sequential crop forwards per slice, encoder_decoder.py:157-203) +
    argmax + the integer confusion matrix, fp32, eval mode."""
    from oracle import stc_oracle as O
    torch.manual_seed(0)
    seg = make_segmentor("stc", num_classes, "fp32")
    bsd = {k: v.detach().clone() for k, v in seg.backbone.state_dict().items()}
    hsd = {k: v.detach().clone() for k, v in seg.decode_head.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    img = torch.rand(slices, 3, 512, 512, generator=g)
    lab = torch.randint(0, num_classes, (slices, 512, 512), generator=g).numpy()
    enc = lambda t: O.head_forward(hsd, O.backbone_forward(bsd, t, False, None), False, None)

    def step():
        with torch.no_grad():
            pred = O.simple_test(O.slide_inference(enc, img, num_classes, (256, 256), (170, 170)))
        return int(O.confusion_matrix(pred.numpy(), lab, num_classes, 255).sum())
    return step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if args.workload == "slide":
        return run_reference_slide(args, cores)
    sample_batch = 1
    step = cpu_step_factory(args.model, args.classes, args.size, sample_batch)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = sample_batch * args.steps / dt
    sample = f"{sample_batch} image(s) of 3x{args.size}x{args.size} per step, fp32, fwd+CE/Dice loss+bwd (no optimizer), torch CPU, {cores} threads"
    line = dict(impl="reference", metric=metric_name(args), value=value, unit="img/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * dt / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", config=dict(workload=workload_name(args), sample=sample),
                cpu_baseline=dict(value=value, unit="img/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit="img/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def run_reference_slide(args, cores):
    step = cpu_slide_factory(args.classes, 1)
    for _ in range(min(args.warmup, 1)):
        step()
    n = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    dt = time.perf_counter() - t0
    value = n / dt
    sample = f"{n} step(s) of ONE 512x512 slice (9 crop forwards of 256x256, stride 170) + argmax + confusion matrix, fp32, torch CPU, {cores} threads"
    line = dict(impl="reference", metric=metric_name(args), value=value, unit="slices/s", n_gpus=args.gpus, steps=n, warmup=min(args.warmup, 1),
                ms_per_step=1e3 * dt / n, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=workload_name(args), sample=sample),
                cpu_baseline=dict(value=value, unit="slices/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit="slices/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def metric_name(args):
    if args.workload == "slide":
        return "STC-UNet sliding-window inference slices/s, 512x512x64 volume, crop 256 / stride 170, bf16, integer confusion matrix"
    if args.model == "stc":
        return METRIC
    return f"{args.model} train img/s @{args.size}x{args.size} {args.dtype} (fwd+loss+bwd+allreduce+Adam)"


def workload_name(args):
    if args.workload == "slide":
        return (f"my_config/STC-UNet.py test_cfg mode='slide' crop 256 stride 170 over a synthetic 512x512x{args.slices} uint8 volume per GPU, "
                f"{args.classes} classes, int64 confusion matrix")
    return (f"{MODEL_NAME[args.model]} fwd+bwd+Adam, batch {args.batch}/GPU of synthetic "
            f"3x{args.size}x{args.size} slices, {args.classes} classes")


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="stc", choices=["stc", "unet", "unet_b", "unetpp"],
                    help="stc = BASELINE.json configs[1-3] (the default line); unet = configs[0]'s model; unetpp = configs[4]; unet_b = mmseg UNet-S5-D16 + FCNHead")
    ap.add_argument("--workload", default="train", choices=["train", "slide"], help="slide = BASELINE.json configs[3]")
    ap.add_argument("--slices", type=int, default=64, help="slide workload: slices per volume (one volume per step)")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the informative PyTorch-eager (autocast bf16, cuDNN/cuBLAS) yardstick")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--classes", type=int, default=3)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="single GPU: launch the step eagerly instead of replaying the captured CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import stc_unet_b200 as S
    from stc_unet_b200 import ops
    from stc_unet_b200.train import Trainer

    if args.workload == "slide":
        return run_slide(args, dev, rank, world, local)
    torch.manual_seed(0)
    seg = make_segmentor(args.model, args.classes, args.dtype).to(dev)
    seg.train()
    trainer = Trainer(seg, lr=1e-5, betas=(0.9, 0.999))
    from stc_unet_b200 import peer as peer_mod
    if world > 1 and trainer.peer is None and peer_mod.peer_exchange_wanted():
        # the multi-GPU line is the peer-memory path; a silent change of exchange would report a different system
        raise SystemExit("bench.py: N > 1 but the NVLink peer-memory exchange could not be set up (symmetric-memory rendezvous failed). "
                         "Set STC_PEER=0 to measure the NCCL exchange deliberately.")

    # synthetic KiTS19-shaped inputs: a small pool of distinct batches, resident in HBM, different data per rank
    g = torch.Generator().manual_seed(1234 + rank)
    pool = 2
    h_img = [torch.rand(args.batch, 3, args.size, args.size, generator=g).pin_memory() for _ in range(pool)]
    h_gt = [torch.randint(0, args.classes, (args.batch, 1, args.size, args.size), generator=g).pin_memory() for _ in range(pool)]
    d_img = [t.to(dev) for t in h_img]
    d_gt = [t.to(dev) for t in h_gt]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    last = {}

    # launches per step, counted on one eager step (a replayed graph launches the same kernels without passing through Python)
    _names = []
    _orig = S._lib.lib.call

    def _tally(name, *a):
        _names.append(name)
        return _orig(name, *a)
    for i in range(3):
        trainer.step(d_img[i % pool], d_gt[i % pool])
    S._lib.lib.call = _tally
    trainer.step(d_img[0], d_gt[0])
    if "call" in S._lib.lib.__dict__:
        del S._lib.lib.__dict__["call"]
    launches_per_step = sum(LAUNCHES.get(n, 1) for n in _names)
    # N > 1: the step is capturable when its exchanges run as our peer-memory kernels (Trainer.peer); with NCCL inside it hung when tried
    use_graph = not args.no_graph and (world == 1 or trainer.peer is not None)
    if use_graph:   # the whole step (fwd + loss + bwd + Adam, ~1000 launches) captured once and replayed; same kernels, no launch gaps
        trainer.capture(d_img[0], d_gt[0])

    def dev_step(i):
        if use_graph:
            last["lv"] = trainer.step_graph(d_img[i % pool], d_gt[i % pool])
        else:
            last["lv"] = trainer.step(d_img[i % pool], d_gt[i % pool])

    for i in range(args.warmup):
        dev_step(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(dev_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    gpu_launches = launches_per_step * args.steps
    ms_per_step = ms_total / args.steps
    value = world * args.batch / (ms_per_step / 1e3)
    loss_val = float(last["lv"]["loss"].detach())

    # ---- end to end: pinned host buffers -> H2D -> step -> D2H of the loss, every step
    e2e = None
    if not args.no_e2e:
        bi = h_img[0].numel() * 4 + h_gt[0].numel() * 8

        class Prefetcher:
            """What a training loader does: the H2D copy of step i+1's batch runs on a copy stream while step i computes.  Every
            step's inputs still travel host -> device inside the timed region; only the waiting is overlapped."""

            def __init__(self, imgs, gts):
                self.imgs, self.gts = imgs, gts
                self.stream = torch.cuda.Stream()
                self.slots = [(torch.empty_like(imgs[0], device=dev), torch.empty_like(gts[0], device=dev)) for _ in range(2)]
                self.ready = [torch.cuda.Event(), torch.cuda.Event()]
                self.consumed = [None, None]
                self.next = None

            def issue(self, i):
                k = i % 2
                if self.consumed[k] is not None:
                    self.stream.wait_event(self.consumed[k])      # the step that last read this slot has taken its copy
                with torch.cuda.stream(self.stream):
                    self.slots[k][0].copy_(self.imgs[i % pool], non_blocking=True)
                    self.slots[k][1].copy_(self.gts[i % pool], non_blocking=True)
                    self.ready[k].record(self.stream)
                self.next = i

            def take(self, i):
                if self.next != i:
                    self.issue(i)
                k = i % 2
                torch.cuda.current_stream().wait_event(self.ready[k])
                return self.slots[k]

            def release(self, i):
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())
                self.consumed[i % 2] = ev

        def make_e2e_step(pf):
            def step(i):
                img, gt = pf.take(i)
                lv = trainer.step_graph(img, gt) if use_graph else trainer.step(img, gt)
                pf.release(i)
                pf.issue(i + 1)                                   # enqueue the next batch's H2D before blocking on this step's result
                last["host_loss"] = float(lv["loss"])            # D2H read of the step's result
            return step
        e2e_step = make_e2e_step(Prefetcher(h_img, h_gt))
        for i in range(2):
            e2e_step(i)
        ms_e2e = timed(e2e_step, args.steps) / args.steps
        e2e = dict(value=world * args.batch / (ms_e2e / 1e3), unit="img/s", h2d_bytes_per_step=bi, d2h_bytes_per_step=4, ms_per_step=ms_e2e,
                   input="reference interface: fp32 NCHW images + int64 labels from pinned host memory")
        # the same step fed the way a GPU-side input pipeline would: decoded uint8 HWC pixels + uint8 labels,
        # normalised / widened on the device; same pixel values as above up to 8-bit quantisation
        u_img = [(t.permute(0, 2, 3, 1) * 255).round().to(torch.uint8).contiguous().pin_memory() for t in h_img]
        u_gt = [t.to(torch.uint8).pin_memory() for t in h_gt]
        (seg.backbone if seg.backbone is not None else seg.decode_head).img_norm_cfg = dict(mean=[0.0], std=[255.0], to_rgb=False)

        if use_graph:
            trainer.capture(u_img[0].to(dev), u_gt[0].to(dev))

        e2e_u8_step = make_e2e_step(Prefetcher(u_img, u_gt))
        for i in range(2):
            e2e_u8_step(i)
        ms_u8 = timed(e2e_u8_step, args.steps) / args.steps
        e2e["uint8_pipeline"] = dict(value=world * args.batch / (ms_u8 / 1e3), unit="img/s", ms_per_step=ms_u8,
                                     h2d_bytes_per_step=u_img[0].numel() + u_gt[0].numel())

    # ---- roofline of the dominant kernel (umma_kernel = tcgen05 implicit GEMM): CUDA events per launch, 2 extra steps
    roofline, breakdown = None, None
    peaks = measured_peaks()
    if not args.no_profile:
        prof = ops.LaunchProfiler(time_dense=True)
        ops.set_profiler(prof)
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        psteps = 2
        for i in range(psteps):
            trainer.step(d_img[i % pool], d_gt[i % pool])   # eager: the per-launch events need the Python-level calls
        t1.record()
        summ = prof.summary()
        ops.set_profiler(None)
        step_ms_prof = t0.elapsed_time(t1) / psteps
        KNAME = {1: "simt", 2: "stc::umma_kernel", 3: "stc::umma_convh_kernel", 4: "stc::umma_wgradh_kernel"}
        per_kernel = {}
        for (kind, eng), v in summ.items():
            d = per_kernel.setdefault(eng, dict(launches=0, flops=0.0, ms=0.0))
            for f in d:
                d[f] += v[f]
        tensor = {e: v for e, v in per_kernel.items() if e >= 2}
        tc_ms = sum(v["ms"] for v in tensor.values())
        if tensor:
            dom = max(tensor, key=lambda e: tensor[e]["ms"])     # the dominant kernel of the step
            dv = tensor[dom]
            achieved = dv["flops"] / (dv["ms"] * 1e-3) / 1e12
            all_achieved = sum(v["flops"] for v in tensor.values()) / (tc_ms * 1e-3) / 1e12
            # dram bytes per launch from the committed `ncu --set full` capture of this kernel (profiles/), if one exists
            traffic, traffic_detail = None, None
            tp = next((q for q in (os.path.join(ROOT, "profiles", n) for n in ("r2_ncu_traffic.json", "r1_ncu_traffic.json")) if os.path.exists(q)), "")
            if os.path.exists(tp):
                traffic_detail = json.load(open(tp)).get(KNAME[dom])
                if traffic_detail:   # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel (the launch is named)
                    traffic = traffic_detail["dram_bytes_read"] + traffic_detail["dram_bytes_write"]
            roofline = dict(bound="tensor", kernel=KNAME[dom], achieved=achieved, peak=peaks["tflops_sustained"], unit="TFLOP/s",
                            frac=achieved / peaks["tflops_sustained"],
                            traffic_source=("committed ncu --set full capture (profiles/), not measured in this run" if traffic is not None else None),
                            peak_source=f"{peaks['source']} (sustained cuBLAS bf16; kernel timed inside a long step)", traffic=traffic,
                            traffic_detail=traffic_detail, launches_per_step=dv["launches"] // psteps, avg_launch_ms=dv["ms"] / max(dv["launches"], 1),
                            algorithmic_tflop_per_step=dv["flops"] / psteps / 1e12, share_of_step=dv["ms"] / psteps / ms_per_step,
                            all_tcgen05_kernels=dict(achieved=all_achieved, frac=all_achieved / peaks["tflops_sustained"],
                                                     share_of_step=tc_ms / psteps / ms_per_step,
                                                     per_kernel={KNAME[e]: dict(launches=v["launches"] // psteps, ms=v["ms"] / psteps,
                                                                                tflops=v["flops"] / (v["ms"] * 1e-3) / 1e12) for e, v in tensor.items()}))
        breakdown = {f"{k[0]}:{KNAME.get(k[1], k[1])}":
                     dict(launches=v["launches"] // psteps, ms=v["ms"] / psteps, tflops=(v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] > 0 else 0.0)
                     for k, v in summ.items()}

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload on the box's host cores
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        step = cpu_step_factory(args.model, args.classes, args.size, 1)
        step()
        t0 = time.perf_counter()
        n = 0
        while n < 1 or (time.perf_counter() - t0 < 10.0 and n < 8):
            step(); n += 1
        dt = (time.perf_counter() - t0) / n
        cpu_baseline = dict(value=1.0 / dt, unit="img/s", cores=cores, kind="port",
                            sample=f"{n} step(s) of 1 image 3x{args.size}x{args.size}, fp32 fwd+loss+bwd, oracle port on torch CPU ({cores} threads)")

    # ---- informative GPU yardstick (rank 0, N=1): the oracle (= the reference's arithmetic in plain PyTorch) under autocast(bf16) on this
    # same B200, cuDNN / cuBLAS eager kernels, fwd + loss + bwd + torch Adam.  Never the target; it answers "is this faster than just
    # running the reference on a B200".
    exchange = "none" if world == 1 else ("nvlink peer-memory kernels (SyncBN stats + gradient all-reduce)" if trainer.peer is not None else "nccl")
    gpu_eager = None
    if rank == 0 and world == 1 and not args.no_gpu_eager:
        try:
            del trainer
            torch.cuda.empty_cache()
            estep = oracle_step_factory(args.model, seg, d_img[0], d_gt[0], autocast_dtype=torch.bfloat16, optimizer=True)
            for _ in range(2):
                estep()
            torch.cuda.synchronize()
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            ne = 5
            for _ in range(ne):
                estep()
            t1.record(); torch.cuda.synchronize()
            ems = t0.elapsed_time(t1) / ne
            gpu_eager = dict(value=args.batch / (ems / 1e3), unit="img/s", ms_per_step=ems,
                             what="oracle modules (plain PyTorch) under torch.autocast(bf16) on this GPU: cuDNN/cuBLAS eager fwd+loss+bwd+Adam, "
                                  f"batch {args.batch}; informative, not the target (BASELINE.md §3)")
        except Exception as e:  # noqa: BLE001 - the yardstick must never take the bench line down
            gpu_eager = dict(unavailable=f"{type(e).__name__}: {str(e)[:200]}")

    if rank == 0:
        fl = FWD_GMAC_PER_IMG[args.model] * 2 * 3 * 1e9   # fwd+bwd algorithmic FLOP per image
        line = dict(metric=metric_name(args), value=value, unit="img/s", n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_per_step,
                    higher_is_better=True, scaling="weak", vs_baseline=None, dtype=args.dtype, data="synthetic",
                    config=dict(workload=workload_name(args), parallelism=f"dp{world}", sync_bn=world > 1, cuda_graph=bool(use_graph),
                                exchange=exchange,
                                cache="per-step working set (tens of GB of activations) >> 126 MB L2; 2 distinct input batches alternate",
                                optimizer="fused Adam lr=1e-5", dropout_ratio=0.1),
                    model_tflops_per_s=value * fl / 1e12, model_frac_of_peak=value * fl / 1e12 / world / peaks["tflops_sustained"],
                    loss=loss_val, roofline=roofline, kernel_breakdown=breakdown, cpu_baseline=cpu_baseline, e2e=e2e,
                    gpu_eager_baseline=gpu_eager, gpu_launches=gpu_launches, clocks=clocks)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs[3]: sliding-window inference over a 512x512x64 volume + integer confusion matrix
# ------------------------------------------------------------------------------------------------
def run_slide(args, dev, rank, world, local):
    """A step = one synthetic 512x512x`slices` uint8 volume per GPU: every slice through test_cfg mode='slide' (crop 256, stride 170:
    9 windows, all windows of a group of slices in ONE forward), argmax, int64 confusion matrix accumulated on the device, one
    all-reduce + 72-byte read-back per volume.  Volumes are independent: ranks shard them with no data-path collective (weak scaling)."""
    import stc_unet_b200 as S
    from stc_unet_b200 import ops
    from stc_unet_b200.metrics import ConfusionMeter
    torch.manual_seed(0)
    seg = make_segmentor("stc", args.classes, args.dtype).to(dev)
    seg.test_cfg = dict(mode="slide", crop_size=(256, 256), stride=(170, 170), max_windows_per_forward=72)
    seg.eval()
    ops.config.cache_eval_weights = True      # a deployed checkpoint: frozen weights, folded / packed operands are kept between forwards
    seg.backbone.img_norm_cfg = dict(mean=[0.0], std=[255.0], to_rgb=False)
    g = torch.Generator().manual_seed(5 + rank)
    group = 8                                    # slices per forward (x 9 windows = 72 crops of 256x256)
    h_vol = torch.randint(0, 256, (args.slices, 512, 512, 3), dtype=torch.uint8, generator=g).pin_memory()
    h_lab = torch.randint(0, args.classes, (args.slices, 512, 512), dtype=torch.uint8, generator=g).pin_memory()
    d_vol, d_lab = h_vol.to(dev), h_lab.to(dev)
    last = {}

    def volume(from_host):
        vol, lab = (h_vol, h_lab) if from_host else (d_vol, d_lab)
        meter = ConfusionMeter(args.classes, 255, device=dev)
        for s0 in range(0, args.slices, group):
            img, lb = vol[s0:s0 + group], lab[s0:s0 + group]
            if from_host:
                img, lb = img.to(dev, non_blocking=True), lb.to(dev, non_blocking=True)
            meter.update(seg.inference_device(img), lb)
        meter.all_reduce()
        last["meter"] = meter
        if from_host:
            last["res"] = meter.compute(["mIoU", "mDice"])     # D2H of the C x C int64 matrix: the volume's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    names, orig = [], S._lib.lib.call
    S._lib.lib.call = lambda name, *a: (names.append(name), orig(name, *a))[1]
    volume(False)
    del S._lib.lib.__dict__["call"]
    launches_per_step = sum(LAUNCHES.get(n, 1) for n in names)
    for _ in range(max(args.warmup - 1, 0)):
        volume(False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(lambda: volume(False), args.steps) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    value = world * args.slices / (ms / 1e3)
    volume(True)
    ms_e2e = timed(lambda: volume(True), args.steps) / args.steps
    res, meter = last["res"], last["meter"]
    peaks = measured_peaks()
    roofline = breakdown = None
    if not args.no_profile:
        prof = ops.LaunchProfiler(time_dense=True)
        ops.set_profiler(prof)
        volume(False)
        summ = prof.summary()
        ops.set_profiler(None)
        KNAME = {1: "simt", 2: "stc::umma_kernel", 3: "stc::umma_convh_kernel", 4: "stc::umma_wgradh_kernel"}
        per = {}
        for (kind, eng), v in summ.items():
            d = per.setdefault(eng, dict(launches=0, flops=0.0, ms=0.0))
            for f in d:
                d[f] += v[f]
        tensor = {e: v for e, v in per.items() if e >= 2}
        if tensor:
            dom = max(tensor, key=lambda e: tensor[e]["ms"])
            dv = tensor[dom]
            ach = dv["flops"] / (dv["ms"] * 1e-3) / 1e12
            roofline = dict(bound="tensor", kernel=KNAME[dom], achieved=ach, peak=peaks["tflops_sustained"], unit="TFLOP/s", frac=ach / peaks["tflops_sustained"],
                            peak_source=f"{peaks['source']} (sustained cuBLAS bf16; kernel timed inside a long step)", traffic=None,
                            launches_per_step=dv["launches"], avg_launch_ms=dv["ms"] / max(dv["launches"], 1),
                            algorithmic_tflop_per_step=dv["flops"] / 1e12, share_of_step=dv["ms"] / ms)
        breakdown = {f"{k[0]}:{KNAME.get(k[1], k[1])}": dict(launches=v["launches"], ms=v["ms"], tflops=(v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] > 0 else 0.0)
                     for k, v in summ.items()}
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        step = cpu_slide_factory(args.classes, 1)
        t0 = time.perf_counter()
        n = 0
        while n < 1 or (time.perf_counter() - t0 < 10.0 and n < 4):
            step(); n += 1
        dt = (time.perf_counter() - t0) / n
        cpu_baseline = dict(value=1.0 / dt, unit="slices/s", cores=cores, kind="port",
                            sample=f"{n} x ONE 512x512 slice (9 crop forwards) + argmax + confusion matrix, fp32, oracle port on torch CPU ({cores} threads)")
    if rank == 0:
        fl = SLIDE_GMAC_PER_SLICE * 2e9
        line = dict(metric=metric_name(args), value=value, unit="slices/s", n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms,
                    higher_is_better=True, scaling="weak", vs_baseline=None, dtype=args.dtype, data="synthetic",
                    config=dict(workload=workload_name(args), parallelism=f"dp{world} (independent volumes, one int64 confusion-matrix all-reduce per volume)",
                                windows_per_slice=9, slices_per_forward=group, eval_bn_folded=True, weights_cached=True,
                                cache="a volume's crops (50 MB uint8) and activations (GBs) >> 126 MB L2"),
                    model_tflops_per_s=value * fl / 1e12, model_frac_of_peak=value * fl / 1e12 / world / peaks["tflops_sustained"],
                    mIoU=res["mIoU"], mDice=res["mDice"], pixels_counted=int(meter.cm.sum()), roofline=roofline, kernel_breakdown=breakdown,
                    cpu_baseline=cpu_baseline,
                    e2e=dict(value=world * args.slices / (ms_e2e / 1e3), unit="slices/s", ms_per_step=ms_e2e,
                             h2d_bytes_per_step=h_vol.numel() + h_lab.numel(), d2h_bytes_per_step=8 * args.classes * args.classes),
                    gpu_launches=launches_per_step * args.steps, clocks=clocks)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _main_with_clean_stdout():
    """stdout must carry exactly ONE JSON line: everything else that libraries print there (e.g. NCCL's version banner)
    is diverted to stderr by pointing fd 1 at fd 2 while the benchmark runs."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    lines = []
    real_print = print

    def capture(*a, **k):
        if k.get("file") in (None, sys.stdout):
            lines.append(" ".join(str(x) for x in a))
        else:
            real_print(*a, **k)
    import builtins
    builtins.print = capture
    try:
        main()
    finally:
        builtins.print = real_print
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    for ln in lines:
        real_print(ln, flush=True)


if __name__ == "__main__":
    _main_with_clean_stdout()
