#!/usr/bin/env python
"""tools/probe/exchange_probe.py -- the n_q-split exchange of BASELINE config 5 (2048^3 fp64, q = 3) over N GPUs, three forms:
NCCL reduce, scatter kernel + library barrier + reduce kernel, ONE kernel with an in-kernel barrier.  Run under torchrun:

    python -m torch.distributed.run --nproc-per-node N tools/probe/exchange_probe.py [--ncu-rank0 OUT.csv]

With --ncu-rank0 rank 0 re-executes itself under `ncu --metrics <nvlink + dram bytes>` restricted to ttv_col_scatter_kernel (the
form without in-kernel waits: ncu replays kernels, and a replayed wait would not see its flags again), so that the bytes the
kernel's stores put on NVLink can be compared with the algorithmic |C| (G-1)/G per GPU."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

METRICS = "gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum"


def nvlink_kib(index):
    """(tx, rx) KiB of user data this GPU has moved over all its NVLink links since the driver loaded (NVML field values
    NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX / _RX, scope = all links), or None when the platform does not expose them"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        vals = pynvml.nvmlDeviceGetFieldValues(h, [(pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, 0xFFFFFFFF),
                                                   (pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, 0xFFFFFFFF)])
        out = []
        for v in vals:
            if v.nvmlReturn != 0:
                return None
            out.append(int(v.value.ullVal))
        return tuple(out)
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ncu-rank0", default="")
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    if args.ncu_rank0 and rank == 0 and not args.child:
        cmd = ["ncu", "--metrics", METRICS, "--clock-control", "none", "-k", "regex:ttv_col_scatter", "-c", "3", "--csv", "--log-file", args.ncu_rank0,
               sys.executable, os.path.abspath(__file__), "--child", "--ncu-rank0", args.ncu_rank0, "--reps", str(args.reps)]
        os.execvp("ncu", cmd)
    import torch
    import torch.distributed as dist
    import ttv_b200
    from ttv_b200.sharded import PeerExchange, make_shard, ttv_sharded
    world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    na, pia, q = [2048, 2048, 2048], [1, 2, 3], 3
    sh = make_shard(q, na, pia, rank, world)
    a = torch.empty(sh.a_count, dtype=torch.float64, device=dev)
    ttv_b200.fill(a, 0x77170001, first=sh.a_offset)
    b = torch.empty(2048, dtype=torch.float64, device=dev)
    ttv_b200.fill(b, 0x77170005)
    c = torch.empty(sh.c_count, dtype=torch.float64, device=dev)
    forms = {"ncclReduce": None, "scatter kernel + library barrier + reduce kernel": PeerExchange(sh.c_count, torch.float64, dev, single_kernel=False)}
    if not args.ncu_rank0:
        forms["ONE kernel, in-kernel barrier"] = PeerExchange(sh.c_count, torch.float64, dev)
    byt = 8 * (2048 ** 3 + 2048 + 2048 ** 2)
    results, nvl = {}, {}
    for name, ex in forms.items():
        for _ in range(3):
            ttv_sharded(q, a, na, pia, b, rank=rank, world=world, c_local=c, reduce_to=0, exchange=ex, asynchronous=True)
        torch.cuda.synchronize()
        nv0 = nvlink_kib(local)               # (NVML can take tens of ms: read it BEFORE the barrier, or the other ranks' timed
        dist.barrier(); torch.cuda.synchronize()   #  windows would include this rank's delay through the exchange itself)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            out, s = ttv_sharded(q, a, na, pia, b, rank=rank, world=world, c_local=c, reduce_to=0, exchange=ex, asynchronous=True)
        e1.record(); torch.cuda.synchronize()
        dist.barrier()
        nv1 = nvlink_kib(local)
        if nv0 is not None and nv1 is not None:
            nvl[name] = ((nv1[0] - nv0[0]) / args.reps / 1024.0, (nv1[1] - nv0[1]) / args.reps / 1024.0)      # MiB per exchange
        t = torch.tensor([e0.elapsed_time(e1) / args.reps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        results[name] = float(t.item())
        if ex is not None and ex.single_kernel and ex.timed_out():
            raise SystemExit(f"rank {rank}: the single-kernel exchange timed out")
    if rank == 0:
        print(f"cfg5 q=3 over {world} GPUs (per GPU: {sh.a_count * 8 / 1e9:.2f} GB of A, partial C {sh.c_count * 8 / 2 ** 20:.0f} MiB, "
              f"algorithmic NVLink bytes out per GPU {sh.c_count * 8 * (world - 1) / world / 2 ** 20:.1f} MiB)")
        for name, ms in results.items():
            link = f"   NVLink user data of rank 0 per exchange (NVML): tx {nvl[name][0]:.1f} MiB, rx {nvl[name][1]:.1f} MiB" if name in nvl else ""
            print(f"  {name:52s} {ms:8.4f} ms   {byt / ms / 1e6:9.1f} GB/s aggregate   {byt / ms / 1e6 / world:8.1f} GB/s per GPU{link}", flush=True)
    del forms, results
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
