"""Per-entry-point time of one training step, measured with CUDA events around EVERY C-ABI call (in-stream, warm clocks; the ncu
launch list under profiles/ is the cold-cache, serialised counterpart).  Usage (GPU box):
    python tools/step_profile.py [--model stc|unet] [--batch 16] [--detail stc_bn_bwd_apply,...] > gpurun_out/step_profile.txt
"""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="stc")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--classes", type=int, default=3)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--detail", default="")
    ap.add_argument("--dense", action="store_true", help="list the dense calls (conv / gemm entry points) grouped by their integer arguments")
    args = ap.parse_args()
    import stc_unet_b200 as S
    from stc_unet_b200 import ops
    from stc_unet_b200.train import Trainer
    # under torchrun (WORLD_SIZE > 1) this profiles rank 0's step of the data-parallel run: SyncBN exchanges + gradient all-reduce included
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    bcfg, hcfg = bench.model_cfg(args.model, args.classes, args.dtype)
    seg = S.EncoderDecoder(bcfg, hcfg).to(dev)
    seg.backbone.init_weights(); seg.decode_head.init_weights()
    seg.train()
    trainer = Trainer(seg, lr=1e-5, betas=(0.9, 0.999))
    img = torch.rand(args.batch, 3, args.size, args.size, device=dev)
    gt = torch.randint(0, args.classes, (args.batch, 1, args.size, args.size), device=dev)
    for _ in range(3):
        trainer.step(img, gt)
    prof = ops.LaunchProfiler(time_all=True)
    ops.set_profiler(prof)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    trainer.step(img, gt)
    t1.record()
    torch.cuda.synchronize()
    ops.set_profiler(None)
    if rank != 0:
        if world > 1:
            dist.barrier()
        return
    agg = collections.OrderedDict()
    per = collections.defaultdict(list)
    for name, nbytes, s, e in prof.all_records:
        ms = s.elapsed_time(e)
        d = agg.setdefault(name, [0, 0.0, 0])
        d[0] += 1; d[1] += ms; d[2] += nbytes
        per[name].append((ms, nbytes))
    total = sum(v[1] for v in agg.values())
    print(f"step (with per-call events) {t0.elapsed_time(t1):.2f} ms; sum over {len(prof.all_records)} C-ABI calls {total:.2f} ms")
    print(f"{'entry point':34s} {'calls':>5s} {'ms':>8s} {'GB touched':>10s} {'GB/s':>8s}")
    for name, (n, ms, nb) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:34s} {n:5d} {ms:8.3f} {nb / 1e9:10.3f} {nb / 1e6 / max(ms, 1e-6):8.0f}")
    if args.dense:
        grp = collections.OrderedDict()
        for (name, nbytes, s, e), sig in zip(prof.all_records, prof.all_sigs):
            if not (name.startswith("stc_conv_") or name.startswith("stc_gemm")):
                continue
            d = grp.setdefault((name, sig), [0, 0.0])
            d[0] += 1; d[1] += s.elapsed_time(e)
        print("--- dense calls by (entry point, integer arguments): calls, total ms")
        for (name, sig), (n, ms) in sorted(grp.items(), key=lambda kv: -kv[1][1]):
            print(f"    {ms:8.3f} ms {n:3d} x {name} {sig}")
    for name in filter(None, args.detail.split(",")):
        print(f"--- {name}")
        for ms, nb in per[name]:
            print(f"    {ms:8.3f} ms {nb / 1e6:10.1f} MB {nb / 1e6 / max(ms, 1e-6):8.0f} GB/s")


if __name__ == "__main__":
    main()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and int(os.environ.get("RANK", "0")) == 0:
        import torch.distributed as dist
        dist.barrier()
