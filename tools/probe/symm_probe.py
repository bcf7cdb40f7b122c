"""probe: does torch symmetric memory give peer-mapped pointers on this box?  torchrun --nproc-per-node 2 tools/probe/symm_probe.py"""
import os
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem
print(rank, "symm_mem api:", [n for n in dir(symm_mem) if not n.startswith("_")][:40], flush=True)
t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "handle:", type(hdl).__name__, [n for n in dir(hdl) if not n.startswith("_")], flush=True)
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "local", hex(t.data_ptr()), flush=True)
t.fill_(float(rank + 1))
hdl.barrier(channel=0)
peer = hdl.get_buffer((rank + 1) % world, (16,), torch.float32)
print(rank, "peer value", peer[:4].tolist(), flush=True)
peer[:4] = 100.0 + rank          # write into the peer's memory
hdl.barrier(channel=0)
torch.cuda.synchronize()
print(rank, "my buffer after peer write", t[:6].tolist(), flush=True)
dist.destroy_process_group()
