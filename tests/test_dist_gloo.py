"""world_size-2 gloo tests (CPU) of the multi-rank host logic: the SyncBN statistic exchange protocol
([sum, sumsq] all-reduce -> global mean / biased var / unbiased running var, identical on every rank), the packed
log-var reduction, and bucketed gradient averaging over the flat arena."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from stc_unet_b200.segmentor import EncoderDecoder
        from stc_unet_b200.train import plan_buckets
        g = torch.Generator().manual_seed(0)
        full = torch.randn(world * 6, 5, 7, 9, generator=g, dtype=torch.float64) * 3 + 10   # |mean| >> std on purpose
        mine = full[rank * 6:(rank + 1) * 6]
        C = 5
        # what ops._bn_forward_stats does: local [sum, sumsq] -> all_reduce -> finalize with count = P * world
        sums = torch.cat([mine.sum(dim=(0, 2, 3)), (mine * mine).sum(dim=(0, 2, 3))])
        dist.all_reduce(sums)
        count = mine.numel() / C * world
        mean = sums[:C] / count
        var = sums[C:] / count - mean * mean
        ok_bn = torch.allclose(mean, full.mean(dim=(0, 2, 3))) and torch.allclose(var, full.var(dim=(0, 2, 3), unbiased=False)) and \
            torch.allclose(var * count / (count - 1), full.var(dim=(0, 2, 3), unbiased=True))
        # backward exchange: [sum g, sum g*xhat] summed over ranks
        gout = torch.randn(world * 6, 5, 7, 9, generator=g, dtype=torch.float64)[rank * 6:(rank + 1) * 6]
        xh = (mine - mean.view(1, C, 1, 1)) / torch.sqrt(var.view(1, C, 1, 1) + 1e-5)
        bs = torch.cat([gout.sum(dim=(0, 2, 3)), (gout * xh).sum(dim=(0, 2, 3))])
        local_dbeta = bs[:C].clone()
        dist.all_reduce(bs)
        gathered = [torch.zeros_like(local_dbeta) for _ in range(world)]
        dist.all_gather(gathered, local_dbeta)
        ok_bwd = torch.allclose(bs[:C], sum(gathered))
        # packed log vars: one all_reduce, mean over ranks
        lv = EncoderDecoder.log_vars_to_host({"decode.loss_bce": torch.tensor(1.0 + rank), "decode.acc_seg": torch.tensor(10.0 * (rank + 1)),
                                              "loss": torch.tensor(2.0 + rank)})
        ok_lv = abs(lv["decode.loss_bce"] - 1.5) < 1e-6 and abs(lv["decode.acc_seg"] - 15.0) < 1e-6 and abs(lv["loss"] - 2.5) < 1e-6
        # bucketed averaging over a flat arena
        sizes = [5, 64, 3, 1000, 7]
        offs, tot = [], 0
        for n in sizes:
            offs.append(tot); tot += (n + 3) // 4 * 4
        flat = torch.full((tot,), float(rank + 1))
        buckets, _ = plan_buckets(list(zip(offs, sizes)), tot, cap_elems=500)
        for (a, b, _) in buckets:
            dist.all_reduce(flat[a:b]); flat[a:b] /= world
        ok_ar = torch.allclose(flat, torch.full((tot,), (1 + world) / 2))
        # peer-exchange agreement: all ranks or none (a CPU "device" can never rendezvous symmetric memory -> try_create gives None)
        from stc_unet_b200 import peer as peer_mod
        cpu = torch.device("cpu")
        ok_peer = peer_mod.agree(object() if rank == 0 else None, cpu) is None and peer_mod.agree(object(), cpu) is not None \
            and peer_mod.try_create(cpu) is None
        q.put((rank, bool(ok_bn), bool(ok_bwd), bool(ok_lv), bool(ok_ar), bool(ok_peer)))
    finally:
        dist.destroy_process_group()


def test_two_rank_protocols_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for r in res:
        assert all(r[1:]), r
