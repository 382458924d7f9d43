"""2-rank data-parallel training step against the oracle on the CONCATENATED batch (launched by tests/test_dist_gpu.py, or by hand:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/dist_worker.py fp32 64

Reference semantics (mmseg/apis/train.py:104-113 + torch nn.SyncBatchNorm): every rank runs the model on its own shard with BN
statistics taken over ALL ranks' pixels, computes its own loss (CE mean over its pixels, Dice over its batch), and the parameter
gradients are averaged over the ranks; running statistics use the global count.  Here rank r holds image r of a batch of `world`
images; the oracle (fp64, on rank 0's GPU) sees the whole batch in one BN and the mean of the per-shard losses.  Our side goes
through Trainer.step: SyncBN statistics via stc_peer_allreduce_small_f64 and the gradient arena via stc_peer_allreduce_arena_f32
(NVLink peer memory; NCCL when STC_PEER=0)."""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


def main():
    dtype, size = sys.argv[1], int(sys.argv[2])
    posbn = len(sys.argv) < 4 or sys.argv[3] != "default"
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    import stc_unet_b200 as S
    from stc_unet_b200.train import Trainer
    from oracle import stc_oracle as O
    from tests.test_model_gpu import build, is_bn_cancelled_bias
    from tests.util import rel_l2

    g = torch.Generator().manual_seed(4242)
    img = torch.rand(world, 3, size, size, generator=g)
    gt = torch.randint(0, 3, (world, 1, size, size), generator=g)
    gt[:, :, :3] = 255
    bb, hd = build(True, 3, dtype, posbn=posbn)          # same seed on every rank -> identical initial weights
    sd_b0 = {k: v.detach().clone() for k, v in bb.state_dict().items()}
    sd_h0 = {k: v.detach().clone() for k, v in hd.state_dict().items()}
    seg = S.EncoderDecoder(bb, hd).to(dev).train()
    assert any(isinstance(m, torch.nn.SyncBatchNorm) for m in seg.modules())
    tr = Trainer(seg, lr=0.0)
    exchange = "peer" if tr.peer is not None else "nccl"
    lv = tr.step(img[rank:rank + 1].to(dev), gt[rank:rank + 1].to(dev))
    torch.cuda.synchronize()
    my_loss = float(lv["loss"])
    losses = [torch.zeros(1, device=dev) for _ in range(world)]
    dist.all_gather(losses, torch.tensor([my_loss], device=dev))
    # every rank must hold bit-identical reduced gradients and running statistics
    flat = tr.arena.flat.clone()
    ref_flat = flat.clone()
    dist.broadcast(ref_flat, 0)
    same_grads = bool(torch.equal(flat, ref_flat))
    bufs = torch.cat([b.detach().float().flatten() for b in seg.buffers()])
    ref_bufs = bufs.clone()
    dist.broadcast(ref_bufs, 0)
    same_bufs = bool(torch.equal(bufs, ref_bufs))
    flags = torch.tensor([float(same_grads), float(same_bufs)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    ok, report = True, {}
    if rank == 0:
        dt64 = torch.float64
        conv = lambda v: v.detach().to(dev).to(dt64) if v.is_floating_point() else v.detach().to(dev)
        bsd = {k: conv(v).requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd_b0.items()}
        hsd = {k: conv(v).requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd_h0.items()}
        nb, nh = {}, {}
        logits = O.head_forward(hsd, O.backbone_forward(bsd, img.to(dev).to(dt64), True, nb), True, nh)
        per = [O.losses(logits[r:r + 1], gt[r:r + 1].to(dev)) for r in range(world)]
        total = sum(p["loss_bce"] + p["loss_dice"] for p in per) / world
        total.backward()
        errs = {}
        for pre, mod, sd in (("b", bb, bsd), ("h", hd, hsd)):
            for k, p in mod.named_parameters():
                if is_bn_cancelled_bias(k):
                    ok &= float(p.grad.abs().max()) <= 1e-3 * max(1.0, float(sd[k].grad.abs().max()))
                    continue
                errs[pre + "." + k] = rel_l2(p.grad, sd[k].grad)
        stat_err = 0.0
        for mod, new in ((bb, nb), (hd, nh)):
            cur = mod.state_dict()
            for k, v in new.items():
                if k.endswith("num_batches_tracked"):
                    ok &= int(cur[k]) == int(v)
                else:
                    stat_err = max(stat_err, rel_l2(cur[k], v))
        loss_err = max(abs(float(losses[r]) - float(per[r]["loss_bce"] + per[r]["loss_dice"])) for r in range(world))
        worst = max(errs, key=errs.get)
        report = dict(dtype=dtype, size=size, world=world, posbn=posbn, exchange=exchange, grads=len(errs), grad_err_median=statistics.median(errs.values()),
                      grad_err_max=errs[worst], grad_err_max_name=worst, running_stat_err=stat_err, loss_err=loss_err,
                      identical_grads_on_all_ranks=bool(flags[0] == 1), identical_buffers_on_all_ranks=bool(flags[1] == 1))
        if dtype == "fp32":
            tol_g, tol_s, tol_l = (1e-4, 1e-5, 1e-5) if posbn else (5e-2, 1e-5, 1e-5)
        else:
            # bf16: the bulk is checked through the median below; the WORST gradient (a BN bias of the last decoder level, whose true
            # gradient nearly cancels) sits at 2.0 here and at 2.6 on one GPU, where the reference's own autocast path is at 4.8
            # (profiles/r2_parity_config2_512_flipfree.json), hence the loose bound on the maximum
            tol_g, tol_s, tol_l = 5.0, 2e-2, 3e-2
        # CoordAtt's conv1 feeds a BN whose sum(dy) vanishes only GLOBALLY under SyncBN, so the single-rank centring trick
        # does not apply per rank: its weight gradient is as ill-conditioned as in torch's own fp32 path (1-3e-4 there) -> 1e-3
        loose = {k: e for k, e in errs.items() if "ca.conv1" in k}
        strict = {k: e for k, e in errs.items() if "ca.conv1" not in k}
        report["grad_err_max_strict"] = max(strict.values())
        report["grad_err_max_ca_conv1"] = max(loose.values()) if loose else 0.0
        ok &= max(strict.values()) <= tol_g and report["grad_err_max_ca_conv1"] <= max(tol_g, 1e-3) and stat_err <= tol_s and loss_err <= tol_l
        if dtype != "fp32":
            # bf16 gradients: the reference's own autocast path is at 0.37 (median, flip-free, single GPU 512x512) of the fp64 oracle and
            # ours at 0.06 (tests/test_model_gpu.py::test_bf16_parity_config2_512, profiles/r2_parity_config2_512_*.json)
            ok &= report["grad_err_median"] <= (0.1 if posbn else 0.6)
        ok &= bool(flags[0] == 1) and bool(flags[1] == 1)
        report["passed"] = bool(ok)
        print("DIST_REPORT " + json.dumps(report), flush=True)
        out_dir = os.path.join(ROOT, "gpurun_out")
        if os.path.isdir(out_dir):
            with open(os.path.join(out_dir, f"dist_report_{dtype}_{size}_{exchange}.json"), "w") as f:
                json.dump(report, f, indent=1)
    verdict = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.broadcast(verdict, 0)
    dist.destroy_process_group()
    sys.exit(0 if float(verdict) == 1.0 else 1)


if __name__ == "__main__":
    main()
