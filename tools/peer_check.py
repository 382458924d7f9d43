"""N >= 2 check of the two NVLink peer-memory collectives (csrc/peer.cu) against NCCL, with their timings.  The 2-rank TRAINING-step
comparison against the oracle (SyncBN + gradient all-reduce through these kernels) is tests/test_dist_gpu.py.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 tools/peer_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import stc_unet_b200 as S
from stc_unet_b200.peer import PeerExchange
px = PeerExchange(dev)
g = torch.Generator(device="cuda").manual_seed(100 + rank)
ok = True
# small fp64 exchanges, back to back (ticket / slot logic), various sizes
for it, n in enumerate([128, 2048, 4096, 6, 1, 128, 128, 512]):
    t = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
    ref = t.clone(); dist.all_reduce(ref)
    px.allreduce_small_(t)
    err = float((t - ref).abs().max())
    ok &= err < 1e-12
    if rank == 0: print(f"small n={n} max err {err:.2e}")
# results must be bit-identical on all ranks
t = torch.randn(1000, dtype=torch.float64, device=dev, generator=g); px.allreduce_small_(t)
gath = [torch.empty_like(t) for _ in range(world)]; dist.all_gather(gath, t)
ok &= all(torch.equal(gath[0], x) for x in gath)
# arena all-reduce: odd sizes, offsets, repeated
total = 41_061_452 // 4 * 4 + 8
arena = px.alloc_arena(total)
for (a, b) in [(0, total), (0, 1000), (4, 4 + 12345), (1024, 1024 + 7), (total - 100004, total), (0, total)]:
    arena.copy_(torch.randn(total, device=dev, generator=g))
    ref = arena[a:b].clone(); dist.all_reduce(ref, op=dist.ReduceOp.AVG)
    before = arena.clone()
    torch.cuda.synchronize(); dist.barrier()
    px.allreduce_arena_(a, b)
    torch.cuda.synchronize()
    err = float((arena[a:b] - ref).abs().max())
    untouched = torch.equal(arena[:a], before[:a]) and torch.equal(arena[b:], before[b:])
    ok &= err < 1e-6 and untouched
    if rank == 0: print(f"arena [{a},{b}) max err {err:.2e} untouched {untouched}")
    dist.barrier()
gath = [torch.empty(1000, device=dev) for _ in range(world)]; dist.all_gather(gath, arena[:1000].contiguous())
ok &= all(torch.equal(gath[0], x) for x in gath)
# timing of the full-arena all-reduce (164 MB) vs NCCL
for name, fn in (("peer", lambda: px.allreduce_arena_(0, total)), ("nccl", lambda: dist.all_reduce(arena, op=dist.ReduceOp.AVG))):
    fn(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(f"{name} all-reduce of {total * 4 / 1e6:.0f} MB: {e0.elapsed_time(e1) / 5:.3f} ms")
    dist.barrier()
for name, fn in (("peer", lambda t: px.allreduce_small_(t)), ("nccl", lambda t: dist.all_reduce(t))):
    t = torch.randn(128, dtype=torch.float64, device=dev)
    fn(t); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): fn(t)
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(f"{name} small all-reduce (128 fp64): {e0.elapsed_time(e1) / 50 * 1e3:.1f} us")
    dist.barrier()
flag = torch.tensor([1.0 if ok else 0.0], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0: print("PEER TEST", "PASSED" if float(flag) == 1.0 else "FAILED")
dist.destroy_process_group()
sys.exit(0 if float(flag) == 1.0 else 1)
