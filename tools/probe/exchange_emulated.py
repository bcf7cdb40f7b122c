#!/usr/bin/env python
"""tools/probe/exchange_emulated.py -- the scatter kernel of the fused n_q-split exchange on ONE GPU, for ncu.

ncu refuses kernels that touch another PROCESS's peer-mapped memory, so the DRAM side of the fused kernel is captured with
the peers emulated: the slab of BASELINE config 5 one GPU holds at N = 8 (outer = 1, n_q = 256, inner = 2048^2, fp64, 8.59 GB)
is contracted by ttv_col_scatter_kernel as "rank 0 of 8", its eight blocks of partial sums going to eight workspaces that
all live on this GPU.  What ncu then shows is the kernel's own traffic: A read once, the partial C (32 MiB) written once --
the stores that travel over NVLink on a real 8-GPU box.

    ncu --set full --clock-control none -k regex:ttv_col_scatter -c 1 -o gpurun_out/scatter python tools/probe/exchange_emulated.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
import ttv_b200  # noqa: E402
from ttv_b200.sharded import PeerExchange  # noqa: E402

world, nq, inner = 8, 256, 2048 * 2048
a = torch.empty(nq * inner, dtype=torch.float64, device="cuda")
ttv_b200.fill(a, 0x77170001)
b = torch.empty(nq, dtype=torch.float64, device="cuda")
ttv_b200.fill(b, 0x77170005)
blk = PeerExchange.block(inner, world)
ws = [torch.zeros(world * blk, dtype=torch.float64, device="cuda") for _ in range(world)]
for _ in range(3):
    ttv_b200.ttv_view_scatter(1, nq, inner, a, b, [w.data_ptr() for w in ws], 0, blk)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ttv_b200.ttv_view_scatter(1, nq, inner, a, b, [w.data_ptr() for w in ws], 0, blk)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
byt = 8 * (nq * inner + nq + inner)
print(f"ttv_col_scatter_kernel, rank 0 of {world} emulated on one GPU: {ms:.4f} ms, {byt / ms / 1e6:.1f} GB/s "
      f"(A {nq * inner * 8 / 1e9:.2f} GB read, partial C {inner * 8 / 2 ** 20:.0f} MiB scattered into {world} workspaces)")
