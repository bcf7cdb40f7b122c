#!/usr/bin/env python
"""tools/probe/numa_probe.py -- under torchrun: where do the ranks' pinned buffers live relative to their GPUs, and what
does that do to H2D bandwidth when all ranks copy at once?
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/probe/numa_probe.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist
from ttv_b200.sharded import bind_host_to_gpu, gpu_numa_cpus

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
GIB = 4
n = GIB * (1 << 30) // 4
d = torch.empty(n, dtype=torch.float32, device=dev)


def h2d_rate(h, together):
    torch.cuda.synchronize()
    if together:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return GIB * 1.073741824 / dt


pr = torch.cuda.get_device_properties(local)
info = {"rank": rank, "pci": f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0",
        "numa_cpus": len(gpu_numa_cpus(local) or []), "affinity_before": len(os.sched_getaffinity(0)), "cpu_count": os.cpu_count()}
try:
    bdf = info["pci"]
    info["numa_node"] = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
except Exception:
    info["numa_node"] = None
h0 = torch.empty(n, dtype=torch.float32, pin_memory=True); h0.fill_(1.0)
h2d_rate(h0, True)
info["unbound_together_gbs"] = round(h2d_rate(h0, True), 1)
for r in range(world):                       # one rank at a time
    if r == rank:
        info["unbound_alone_gbs"] = round(h2d_rate(h0, False), 1)
    dist.barrier()
del h0
info["bound"] = bind_host_to_gpu(local)
info["affinity_after"] = len(os.sched_getaffinity(0))
h1 = torch.empty(n, dtype=torch.float32, pin_memory=True); h1.fill_(1.0)
h2d_rate(h1, True)
info["bound_together_gbs"] = round(h2d_rate(h1, True), 1)
for r in range(world):
    if r == rank:
        info["bound_alone_gbs"] = round(h2d_rate(h1, False), 1)
    dist.barrier()
out = [None] * world
dist.all_gather_object(out, info)
if rank == 0:
    for o in out:
        print(json.dumps(o))
    print("aggregate unbound together %.1f GB/s, bound together %.1f GB/s" % (sum(o["unbound_together_gbs"] for o in out), sum(o["bound_together_gbs"] for o in out)))
dist.destroy_process_group()
