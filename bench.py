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
FWD_GMAC_PER_IMG = {"stc": 513.86, "unet": 123.70}   # SURVEY.md §8(d), forward MACs per 512x512 image
LOSS_CFG = [dict(type="CrossEntropyLoss", use_sigmoid=False, loss_name="loss_bce", loss_weight=1.0),
            dict(type="DiceLoss", loss_name="loss_dice", loss_weight=1.0)]
# kernel launches per C-ABI call (default 1); memsets are not counted
LAUNCHES = {"stc_bn_reduce": 2, "stc_bn_bwd_reduce": 2, "stc_seg_loss_fwd": 2, "stc_upcat_bwd": 2}


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
    from oracle import stc_oracle as O
    import stc_unet_b200 as S
    torch.manual_seed(0)
    bcfg, hcfg = model_cfg(kind, num_classes, "fp32")
    bb, hd = S.build_backbone(bcfg), S.build_head(dict(hcfg, dropout_ratio=0.0))
    bb.init_weights(); hd.init_weights()
    bsd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in bb.state_dict().items()}
    hsd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in hd.state_dict().items()}
    g = torch.Generator().manual_seed(0)
    img = torch.rand(batch, 3, size, size, generator=g)
    gt = torch.randint(0, num_classes, (batch, 1, size, size), generator=g)

    def step():
        for d in (bsd, hsd):
            for v in d.values():
                v.grad = None
        out = O.forward_train(bsd, hsd, img, gt, True, {}, {})
        (out["loss_bce"] + out["loss_dice"]).backward()
        return float(out["loss_bce"])
    return step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
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
    line = dict(impl="reference", metric=METRIC, value=value, unit="img/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * dt / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", config=dict(workload=workload_name(args), sample=sample),
                cpu_baseline=dict(value=value, unit="img/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit="img/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def workload_name(args):
    return (f"my_config/{'STC-UNet' if args.model == 'stc' else 'U-Net'}.py fwd+bwd+Adam, batch {args.batch}/GPU of synthetic "
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
    ap.add_argument("--model", default="stc", choices=["stc", "unet"])
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

    torch.manual_seed(0)
    bcfg, hcfg = model_cfg(args.model, args.classes, args.dtype)
    seg = S.EncoderDecoder(bcfg, hcfg).to(dev)
    seg.backbone.init_weights(); seg.decode_head.init_weights()
    seg.train()
    trainer = Trainer(seg, lr=1e-5, betas=(0.9, 0.999))

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
        seg.backbone.img_norm_cfg = dict(mean=[0.0], std=[255.0], to_rgb=False)

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
            tp = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")
            if os.path.exists(tp):
                traffic_detail = json.load(open(tp)).get(KNAME[dom])
                if traffic_detail:   # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel (the launch is named)
                    traffic = traffic_detail["dram_bytes_read"] + traffic_detail["dram_bytes_write"]
            roofline = dict(bound="tensor", kernel=KNAME[dom], achieved=achieved, peak=peaks["tflops_sustained"], unit="TFLOP/s",
                            frac=achieved / peaks["tflops_sustained"],
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

    if rank == 0:
        fl = FWD_GMAC_PER_IMG[args.model] * 2 * 3 * 1e9   # fwd+bwd algorithmic FLOP per image
        line = dict(metric=METRIC, value=value, unit="img/s", n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_per_step,
                    higher_is_better=True, scaling="weak", vs_baseline=None, dtype=args.dtype, data="synthetic",
                    config=dict(workload=workload_name(args), parallelism=f"dp{world}", sync_bn=world > 1, cuda_graph=bool(use_graph),
                                exchange=("none" if world == 1 else ("nvlink peer-memory kernels (SyncBN stats + gradient all-reduce)" if trainer.peer is not None
                                                                      else "nccl")),
                                cache="per-step working set (tens of GB of activations) >> 126 MB L2; 2 distinct input batches alternate",
                                optimizer="fused Adam lr=1e-5", dropout_ratio=0.1),
                    model_tflops_per_s=value * fl / 1e12, model_frac_of_peak=value * fl / 1e12 / world / peaks["tflops_sustained"],
                    loss=loss_val, roofline=roofline, kernel_breakdown=breakdown, cpu_baseline=cpu_baseline, e2e=e2e,
                    gpu_launches=gpu_launches, clocks=clocks)
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
