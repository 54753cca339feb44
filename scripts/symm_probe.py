import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = symm_mem.empty((1024, 64), dtype=torch.uint8, device="cuda")
t.fill_(rank + 1)
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast", hdl.has_multicast_support, flush=True)
hdl.barrier()
peer = (rank + 1) % world
pt = hdl.get_buffer(peer, (1024, 64), torch.uint8)
x = pt.clone()
torch.cuda.synchronize()
print(rank, "peer value", int(x[0, 0]), "expected", peer + 1, flush=True)
hdl.barrier()
dist.destroy_process_group()
