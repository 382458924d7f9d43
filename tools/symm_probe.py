"""Probe (N>=2): does torch.distributed._symmetric_memory rendezvous work on this box, and what does a peer pointer look like?"""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = symm.empty(4096, dtype=torch.float64, device=torch.device("cuda", local))
t.fill_(rank + 1)
hdl = symm.rendezvous(t, dist.group.WORLD.group_name if hasattr(dist.group.WORLD, "group_name") else dist.group.WORLD)
print(rank, "rendezvous ok", type(hdl).__name__, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal_pad_ptrs", [hex(p) for p in hdl.signal_pad_ptrs],
      "world", hdl.world_size, "rank", hdl.rank, flush=True)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (8,), torch.float64)
print(rank, "peer view", peer.tolist(), flush=True)
dist.destroy_process_group()
