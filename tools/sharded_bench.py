#!/usr/bin/env python
"""tools/sharded_bench.py -- BASELINE config 5: order-3 fp64 n=(2048,2048,2048) (68.7 GB) sharded over the GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/sharded_bench.py [--extent 2048]

The global tensor is cut along its slowest mode (mode 3, first-order layout): q = 1, 2 are free-mode splits without
communication, q = 3 contracts the split mode (n_q split) and finishes with ONE NCCL reduce of the 32 MiB partial C
(ttv_b200/sharded.py).  Strong scaling: the global size is fixed, every rank holds 1/N of it.  Times are CUDA events on
the device, max over ranks, reduce included; sampled outputs are checked against a host long-double dot on regenerated
data.  One JSON line per q on rank 0.
"""
import argparse
import json
import os
import statistics
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import ttv_b200  # noqa: E402
from ttv_b200.sharded import make_shard, ttv_sharded  # noqa: E402

SEED_A, SEED_B = 0x77170001, 0x77170002



def synth_f64(seed: int, idx):
    """the synthetic generator of the fill kernel (csrc/numeric.cuh: splitmix64 of seed ^ j, top 53 bits mapped to
    [-1, 1)) restated in numpy for the sampled self-checks"""
    with np.errstate(over="ignore"):
        z = (np.asarray(idx, dtype=np.uint64) ^ np.uint64(seed)) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--extent", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--fused", action="store_true", help="also time q=3 with the exchange fused into the kernel (PeerExchange)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.extent
    na, pia = [n, n, n], [1, 2, 3]
    sh = make_shard(1, na, pia, rank, world)
    a = torch.empty(sh.a_count, dtype=torch.float64, device=dev)
    ttv_b200.fill(a, SEED_A, first=sh.a_offset)
    peak = 6553.9
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    for q in (1, 2, 3):
        shq = make_shard(q, na, pia, rank, world)
        b = torch.empty(n, dtype=torch.float64, device=dev)
        ttv_b200.fill(b, SEED_B + q)
        c = torch.full((shq.c_count,), float("nan"), dtype=torch.float64, device=dev)

        def step():
            ttv_sharded(q, a, na, pia, b, rank=rank, world=world, c_local=c, reduce_to=0)

        for _ in range(3):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.reps)]
        for e0, e1 in evs:
            e0.record(); step(); e1.record()
        torch.cuda.synchronize()
        ms = statistics.median(e0.elapsed_time(e1) for e0, e1 in evs)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        # sampled check (rank 0 holds the reduced C for q = 3; every rank checks its own slab otherwise)
        if not (shq.kind == "nq" and rank != 0):
            rng = np.random.default_rng(7 + rank)
            inner = n ** (q - 1)
            bh = b.cpu().numpy().astype(np.longdouble)
            for j in rng.integers(0, shq.c_count, 16):
                jg = int(j) + shq.c_offset
                o, i = divmod(jg, inner)
                idx = (o * n + np.arange(n)) * inner + i
                fiber = synth_f64(SEED_A, idx).astype(np.longdouble)
                want = float(np.dot(fiber, bh))
                tol = 2 * n * (np.finfo(np.float64).eps / 2) * float(np.dot(np.abs(fiber), np.abs(bh))) + 1e-300
                got = float(c[int(j)].item())
                assert abs(got - want) <= tol, (q, jg, got, want, tol)
        if rank == 0:
            byt = 8 * (n ** 3 + n + n ** 2)
            print(json.dumps({"config": f"cfg5 fp64 n=({n},{n},{n}) first-order", "q": q, "n_gpus": world,
                              "split": shq.kind + ("" if shq.kind == "free" else " + NCCL reduce"), "ms": round(ms, 4),
                              "gbs_aggregate": round(byt / ms / 1e6, 1), "gbs_per_gpu": round(byt / ms / 1e6 / world, 1),
                              "frac_of_measured_peak_per_gpu": round(byt / ms / 1e6 / world / peak, 3),
                              "gflops": round(2 * n ** 3 / ms / 1e6, 1)}), flush=True)
    if args.fused and world > 1:
        from ttv_b200.sharded import PeerExchange
        q = 3
        ex = PeerExchange(n * n, torch.float64, dev)
        b = torch.empty(n, dtype=torch.float64, device=dev)
        ttv_b200.fill(b, SEED_B + q)
        out = {}

        def step():
            out["c"], out["sh"] = ttv_sharded(q, a, na, pia, b, rank=rank, world=world, exchange=ex)

        for _ in range(3):
            step()
        dist.barrier(); torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.reps)]
        for e0, e1 in evs:
            e0.record(); step(); e1.record()
        torch.cuda.synchronize()
        ms = statistics.median(e0.elapsed_time(e1) for e0, e1 in evs)
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        # every rank checks samples of ITS block of C
        rng = np.random.default_rng(17 + rank)
        c, shq = out["c"], out["sh"]
        inner = n * n
        bh = b.cpu().numpy().astype(np.longdouble)
        for j in rng.integers(0, shq.c_count, 8):
            i = int(j) + shq.c_offset
            idx = np.arange(n) * inner + i
            fiber = synth_f64(SEED_A, idx).astype(np.longdouble)
            want = float(np.dot(fiber, bh))
            tol = 2 * n * (np.finfo(np.float64).eps / 2) * float(np.dot(np.abs(fiber), np.abs(bh))) + 1e-300
            assert abs(float(c[int(j)].item()) - want) <= tol, ("fused", i)
        if rank == 0:
            byt = 8 * (n ** 3 + n + n ** 2)
            print(json.dumps({"config": f"cfg5 fp64 n=({n},{n},{n}) first-order", "q": q, "n_gpus": world,
                              "split": "nq, exchange fused into the kernel's stores over NVLink peer memory (reduce-scatter)",
                              "ms": round(ms, 4), "gbs_aggregate": round(byt / ms / 1e6, 1), "gbs_per_gpu": round(byt / ms / 1e6 / world, 1),
                              "frac_of_measured_peak_per_gpu": round(byt / ms / 1e6 / world / peak, 3)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
