#!/usr/bin/env python
"""bench.py -- the TTV hot path on B200, measured against the HBM roofline, with the reference's CPU path beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W]          own arm (CUDA kernels through the C-ABI)
    python bench.py --impl reference ...                         the reference's own CPU implementation (oracle/_ref)
    torchrun --nproc-per-node N bench.py --gpus N ...            one rank per GPU (NCCL)

Workload (BASELINE.json configs[1], symmetric sweep): an order-4 fp32 tensor with a 256^4 slab (16 GiB) PER GPU,
first-order layout, and one STEP = the four products q = 1, 2, 3, 4 on it.  With N ranks the global tensor is
(256, 256, 256, 256*N) sharded along its slowest mode (weak scaling): q = 1..3 are free-mode splits without
communication, q = 4 contracts the split mode, so every rank reduces over its 256 rows and the partial C (64 MiB) is
summed with one NCCL reduce.  N = 1 is exactly the named 256^4 case.

metric = effective HBM GB/s = algorithmic bytes / time, algorithmic bytes per product = 4 * (N_el + n_q + N_el/n_q)
(read A once, read b once, write C once).  Inputs (16 GiB) are far larger than L2 (126 MB), so no flush is
needed between iterations.  `value` has the inputs resident in HBM; `e2e` is the same step through the public API
with HOST (pinned) buffers, host<->device copies inside the timed region.

Beside the contract's keys the line carries (none of them inside the timed region of `value`):
  configs   N = 1: EVERY named BASELINE config (ttv_b200/workloads.py "named": cfg1, the symmetric fp32/fp64 sweep, the
            asymmetric fp32/int32 sweep, the complex family with last-order / random layouts, the cfg5 slab: 254 products)
            timed device-resident AND checked on sampled fibers against a host long-double dot (bit-exact for int32):
            n, min / median GB/s, everything below 0.8 x 8 TB/s, parity failures, the whole table
  cfg5      BASELINE configs[4]: 2048^3 fp64 STRONG scaling over the N ranks, q = 1 (free split) and q = 3 (n_q split +
            fused exchange), per-rank times, checked the same way
  multi_gpu_selfcheck   N > 1: small integer tensors through every exchange form against numpy, bit for bit
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

EXT = 256                       # extent of every mode of the per-GPU slab
ORDER = 4
DTYPE = "f32"
ELEM = 4
SEED_A, SEED_B = 0x77170001, 0x77170002
METRIC = "TTV effective HBM GB/s"
# how oracle/_ref is compiled (oracle/Makefile REFFLAGS): built in the build container and shipped to the GPU box, whose CPU
# is not known there, so it targets x86-64-v3 (AVX2 + FMA) instead of BASELINE.md's -march=native; OpenBLAS picks its
# kernels at run time (DYNAMIC_ARCH), so the gemv inside is unaffected
REF_BUILD = "g++ -O3 -march=x86-64-v3 -fopenmp -DNDEBUG, not -march=native: built off the box"


def algo_bytes(na, q, elem=ELEM):
    n = int(np.prod(na, dtype=object))
    return elem * (n + na[q - 1] + n // na[q - 1])


def workload_name(n, nccl_reduce=False):
    how = "n_q split + NCCL reduce" if nccl_reduce else "n_q split, the whole exchange in the product's kernel: stores over NVLink peer memory, in-kernel barrier, slot sum"
    return (f"cfg2 symmetric sweep: order-4 fp32 n=(256,256,256,{EXT * n}) first-order, one step = q=1..4; "
            f"{'single GPU' if n == 1 else f'sharded along mode 4 over {n} GPUs (q=1..3 free split, q=4 {how})'}")


# ---------------------------------------------------------------------------------------------------------------------
# clocks during the timed region (NVML)
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int, period: float = 0.02):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self.index, self.period = index, period

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception:
            self.nv = None
        return self

    def _run(self):
        nv = self.nv
        names = {getattr(nv, k): k for k in dir(nv) if k.startswith("nvmlClocksEventReason") or k.startswith("nvmlClocksThrottleReason")}
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((mhz, util))
                for bit, name in names.items():
                    if isinstance(bit, int) and bit and (mask & bit) == bit and bit & (bit - 1) == 0:
                        self.reasons.add(name.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", ""))
            except Exception:
                pass
            time.sleep(self.period)

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=1)

    def summary(self):
        busy = [m for m, u in self.samples if u > 0] or [m for m, _ in self.samples]
        reasons = sorted(r for r in self.reasons if r not in ("None", "GpuIdle", "ApplicationsClocksSetting"))
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------------------
# the reference's CPU implementation (oracle/_ref): cpu_baseline leg and --impl reference
# ---------------------------------------------------------------------------------------------------------------------
def load_reference():
    from oracle.oracle import Oracle, Reference
    helpers = Oracle()
    for blas in (True, False):
        try:
            if Reference.available(blas=blas):
                return Reference(blas=blas), helpers, "reference"
        except OSError:
            continue
    return None, helpers, "port"


def cpu_step(ref, helpers, kind, a, na, pia, bs, cs):
    """one step (q = 1..4) on the host; returns seconds"""
    t0 = time.perf_counter()
    for q in range(1, ORDER + 1):
        if kind == "reference":
            ref.ttv(q, a, na, pia, bs[q - 1], combo=("par_loop", "subtensor", "all"), helpers=helpers, c0=cs[q - 1])
        else:
            cs[q - 1][:] = helpers.ttv(q, a, na, pia, bs[q - 1])
    return time.perf_counter() - t0


def cpu_measure(last_extent, steps, warmup, budget_s=None, a=None):
    ref, helpers, kind = load_reference()
    na = [EXT, EXT, EXT, last_extent]
    pia = [1, 2, 3, 4]
    n = int(np.prod(na))
    if a is None:
        a = np.empty(n, np.float32)
        chunk = 1 << 24
        pattern = helpers.fill(DTYPE, chunk, SEED_A)
        for s in range(0, n, chunk):
            a[s:s + chunk] = pattern[: min(chunk, n - s)]
    a = a[:n]
    bs = [helpers.fill(DTYPE, na[q - 1], SEED_B + q) for q in range(1, ORDER + 1)]
    cs = [np.zeros(n // na[q - 1], np.float32) for q in range(1, ORDER + 1)]
    for _ in range(warmup):
        cpu_step(ref, helpers, kind, a, na, pia, bs, cs)
    times = []
    t_start = time.perf_counter()
    for _ in range(steps):
        for c in cs:
            c.fill(0)                       # the non-BLAS column kernel accumulates (matrix_times_vector.h:124)
        times.append(cpu_step(ref, helpers, kind, a, na, pia, bs, cs))
        if budget_s is not None and time.perf_counter() - t_start > budget_s:
            break
    total_bytes = sum(algo_bytes(na, q) for q in range(1, ORDER + 1))
    cores = ref.cores() if ref is not None else 1
    blas = bool(ref is not None and ref.blas)
    return {"seconds_per_step": statistics.mean(times), "steps_done": len(times), "bytes_per_step": total_bytes,
            "gbs": total_bytes / statistics.mean(times) / 1e9, "kind": kind, "cores": cores, "blas": blas, "na": na}


def cpu_measure_shape(dt, na, pia, qs, combo, steps, warmup, budget_s):
    """the reference's CPU path on one more shape / policy (cpu_baseline_asym): seconds per step = all q of qs"""
    ref, helpers, kind = load_reference()
    npdt = {"f32": np.float32, "i32": np.int32, "f64": np.float64}[dt]
    n = int(np.prod(na))
    a = np.empty(n, npdt)
    chunk = 1 << 24
    pattern = helpers.fill(dt, chunk, SEED_A)
    for s0 in range(0, n, chunk):
        a[s0:s0 + chunk] = pattern[: min(chunk, n - s0)]
    bs = {q: helpers.fill(dt, na[q - 1], SEED_B + q) for q in qs}
    cs = {q: np.zeros(n // na[q - 1], npdt) for q in qs}
    times = []
    t_start = time.perf_counter()
    for it in range(warmup + steps):
        for c in cs.values():
            c.fill(0)
        t0 = time.perf_counter()
        for q in qs:
            if kind == "reference":
                ref.ttv(q, a, na, pia, bs[q], combo=combo, helpers=helpers, c0=cs[q])
            else:
                cs[q][:] = helpers.ttv(q, a, na, pia, bs[q], slicing=combo[1])
        if it >= warmup:
            times.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_start > budget_s and times:
            break
    elem = np.dtype(npdt).itemsize
    total_bytes = sum(algo_bytes(na, q, elem) for q in qs)
    return {"gbs": total_bytes / statistics.mean(times) / 1e9, "steps_done": len(times), "kind": kind,
            "cores": ref.cores() if ref is not None else 1, "blas": bool(ref is not None and ref.blas)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # size the sample so that (steps + warmup) steps end within ~150 s: probe on a small slab first
    probe = cpu_measure(8, 1, 1)
    rate = probe["gbs"] * 1e9
    want_s = 150.0 / max(1, args.steps + args.warmup)
    last = EXT
    try:
        avail = int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) * 1024
    except Exception:
        avail = 32 << 30
    while last > 8 and (4 * algo_bytes([EXT, EXT, EXT, last], 1) / rate > want_s or 4 * EXT ** 3 * last * 2.5 > avail):
        last //= 2
    m = cpu_measure(last, args.steps, args.warmup)
    sample = (f"unmodified reference headers ({'OpenBLAS ' if m['blas'] else 'non-BLAS '}par_loop/subtensor/all, OpenMP; {REF_BUILD}), "
              f"n=({EXT},{EXT},{EXT},{last}) fp32 {'= the full per-GPU tensor' if last == EXT else 'slab of the 256^4 tensor'}, q=1..4 per step")
    line = {"impl": "reference", "metric": METRIC, "value": round(m["gbs"], 2), "unit": "GB/s", "n_gpus": args.gpus,
            "steps": m["steps_done"], "warmup": args.warmup, "ms_per_step": round(m["seconds_per_step"] * 1e3, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": {"workload": workload_name(args.gpus), "layout": "first-order", "per_gpu_tensor_bytes": EXT ** 4 * ELEM,
                       "timed_on": "host CPU cores",
                       # what this arm actually times: a rate (GB/s) on a bounded sample of the PER-GPU tensor, not the N x tensor
                       "sampled_shape": [EXT, EXT, EXT, last], "sampled_bytes_per_step": m["bytes_per_step"],
                       "sample_is": "the whole per-GPU 256^4 tensor" if last == EXT else "a slab of the per-GPU 256^4 tensor along mode 4",
                       "l2": "inputs are larger than the CPU's caches; no flush needed",
                       "timing": "wall clock around every step (q = 1..4), mean over the steps"},
            "cpu_baseline": {"value": round(m["gbs"], 2), "unit": "GB/s", "cores": m["cores"], "kind": m["kind"], "sample": sample},
            "e2e": {"value": round(m["gbs"], 2), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gflops": round(2 * EXT ** 3 * last * 4 / m["seconds_per_step"] / 1e9, 2)}
    emit(line)


# ---------------------------------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------------------------------
def run_own_arm(args):
    import torch
    import torch.distributed as dist
    import ttv_b200
    from ttv_b200.sharded import make_shard, ttv_sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N with N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_bound = False
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        if not args.no_numa_bind:
            # one process per GPU: keep this rank (and the pinned staging buffers it allocates) on the socket of its GPU
            from ttv_b200.sharded import bind_host_to_gpu
            numa_bound = bind_host_to_gpu(local)
        dist.init_process_group("nccl", device_id=dev)

    na_global = [EXT, EXT, EXT, EXT * world]
    pia = [1, 2, 3, 4]
    shards = {q: make_shard(q, na_global, pia, rank, world) for q in range(1, ORDER + 1)}
    sh = shards[1]
    a = torch.empty(sh.a_count, dtype=torch.float32, device=dev)
    ttv_b200.fill(a, SEED_A, first=sh.a_offset)
    bs = {}
    for q in range(1, ORDER + 1):
        bs[q] = torch.empty(na_global[q - 1], dtype=torch.float32, device=dev)
        ttv_b200.fill(bs[q], SEED_B + q)
    cs = {q: torch.full((shards[q].c_count,), float("nan"), dtype=torch.float32, device=dev) for q in range(1, ORDER + 1)}

    # N > 1: the n_q-split product (q = 4) exchanges its partials through the kernel's own stores into the owners' memory
    # (NVLink peer memory, ttv_b200.sharded.PeerExchange) and leaves C distributed like the free splits do; --nccl-reduce
    # takes the plain kernel + ncclReduce-to-rank-0 instead.
    exchange = None
    if world > 1 and not args.nccl_reduce:
        from ttv_b200.sharded import PeerExchange
        try:
            exchange = PeerExchange(shards[ORDER].c_count, torch.float32, dev)
        except Exception as exc:                      # no symmetric memory on this box: plain kernel + ncclReduce
            print(f"bench.py: rank {rank}: peer-memory exchange unavailable ({exc}); using ncclReduce", file=sys.stderr)
        ok = torch.tensor([1 if exchange is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            exchange = None
            args.nccl_reduce = True
    live = {}          # what the last step produced: q -> (C of this rank, its shard)

    def step():
        for q in range(1, ORDER + 1):
            live[q] = ttv_sharded(q, a, na_global, pia, bs[q], rank=rank, world=world, c_local=cs[q], reduce_to=0, exchange=exchange,
                                  asynchronous=True)          # enqueue only: no host round trip between the four products

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- self-check of every product on sampled fibers against a host long-double dot (not the thing measured) ----
    step()
    torch.cuda.synchronize()
    verify_sample(torch, {q: live[q][0] for q in live}, {q: live[q][1] for q in live}, na_global, bs, rank, world)

    total_bytes = sum(algo_bytes(na_global, q) for q in range(1, ORDER + 1))
    total_flops = 2 * int(np.prod(na_global, dtype=object)) * ORDER

    with ClockSampler(local) as clocks:
        for _ in range(max(args.warmup, 3)):
            step()
        barrier()
        launches0 = ttv_b200.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
        launches = ttv_b200.launch_count() - launches0
        ms = e0.elapsed_time(e1)
        if exchange is not None and exchange.single_kernel and exchange.timed_out():
            raise SystemExit(f"bench.py: rank {rank}: the one-kernel exchange gave up waiting for another GPU inside the timed region")
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            lt = torch.tensor([launches], device=dev, dtype=torch.int64)
            dist.all_reduce(lt)
            launches = int(lt.item())

        # ---- per-product timing of the dominant kernel (events around every launch on the launching stream) ----
        per_q = {}
        for q in range(1, ORDER + 1):
            reps = max(5, min(args.steps, 20))
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
            barrier()
            for s, e in evs:
                s.record()
                local_product(ttv_b200, q, a, shards[q], pia, bs[q], cs[q])
                e.record()
            torch.cuda.synchronize()
            per_q[q] = statistics.mean(s.elapsed_time(e) for s, e in evs)
    clk = clocks.summary()

    ms_per_step = ms / args.steps
    value = total_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel: the column-GEMV kernel (3 of the 4 launches of a step) -------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    col_q = [2, 3, 4]
    col_bytes = [algo_bytes(list(shards[q].na_local), q) for q in col_q]
    col_ms = [per_q[q] for q in col_q]
    achieved = sum(col_bytes) / (sum(col_ms) * 1e-3) / 1e9
    # the kernel the chooser actually launches for these products (ttv_b200_plan: the same code path the launch takes), and
    # the ncu DRAM traffic recorded for exactly that kernel -- a chooser change makes the labels differ and traffic null
    ctype = {"f32": "float", "f64": "double"}[DTYPE]
    labels = []
    for q in col_q:
        pl = ttv_b200.plan(q, list(shards[q].na_local), pia, dtype=DTYPE)
        fam = {1: "dot", 2: "col", 3: "stream", 4: "colx", 5: "dotf"}.get(pl["kernel"], "?")
        labels.append(f"ttv_{fam}_kernel<{ctype},{pl['vec']},{pl['nu']},{pl['ku']}>")
    kernel_name = max(set(labels), key=labels.count)
    traffic, traffic_note = None, None
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        if rec.get("kernel") == kernel_name and all(l == kernel_name for l in labels):
            traffic = rec.get("col_kernel_dram_bytes_per_launch")
            traffic_note = rec.get("source")
        else:
            traffic_note = f"profiles/roofline_traffic.json was captured for {rec.get('kernel')}, this run launched {sorted(set(labels))}"
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kernel_name, "kernel_per_product": dict(zip([f"q{q}" for q in col_q], labels)),
                "achieved": round(achieved, 1), "peak": peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                "unit": "GB/s", "frac": round(achieved / peak, 4), "frac_of_nominal_8000": round(achieved / 8000.0, 4),
                "traffic": traffic, "traffic_source": traffic_note, "algorithmic_bytes_per_launch": int(statistics.mean(col_bytes)),
                "per_product": {f"q{q}": {"ms": round(per_q[q], 4),
                                          "gbs": round(algo_bytes(list(shards[q].na_local), q) / (per_q[q] * 1e-3) / 1e9, 1)}
                                for q in range(1, ORDER + 1)}}

    # ---- end to end through the public API with host buffers --------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(torch, dist, ttv_b200, args, a, bs, cs, shards, na_global, pia, rank, world, dev, total_bytes, exchange)

    cpu, cpu_asym = None, None
    if rank == 0 and world == 1 and not args.no_cpu:
        a_np = None
        try:
            # the host copy of the tensor the e2e leg just used (16 GiB, already resident): the whole workload, a few steps
            a_np = e2e.pop("_a_host", None) if e2e else None
            last = EXT if a_np is not None else 32
            m = cpu_measure(last, 5, 1, budget_s=25.0, a=a_np)
            cpu = {"value": round(m["gbs"], 2), "unit": "GB/s", "cores": m["cores"], "kind": m["kind"],
                   "sample": (f"{'unmodified reference headers' if m['kind'] == 'reference' else 'oracle port'} "
                              f"({'OpenBLAS ' if m['blas'] else 'non-BLAS '}par_loop/subtensor/all; {REF_BUILD}), "
                              f"{'the whole' if last == EXT else f'slab n=(256,256,256,{last}) fp32 of the'} 256^4 tensor, q=1..4, "
                              f"{m['steps_done']} steps after 1 warm-up")}
        except Exception as exc:  # the baseline is reported, never required for the GPU number
            cpu = {"value": None, "unit": "GB/s", "cores": 0, "kind": "port", "sample": f"failed: {exc}"}
        del a_np
        # the asymmetric family with the policy the reference's README states for it, (par_loop, slice, all)
        # (/root/reference README.md:79, detail/tensor_times_vector.h:706): one member of the sweep, every q
        try:
            na_asym = [16, 1024, 4, 1 << 14]
            m2 = cpu_measure_shape("f32", na_asym, [1, 2, 3, 4], [1, 2, 3, 4], ("par_loop", "slice", "all"), 3, 1, 12.0)
            cpu_asym = {"value": round(m2["gbs"], 2), "unit": "GB/s", "cores": m2["cores"], "kind": m2["kind"],
                        "sample": (f"{'unmodified reference headers' if m2['kind'] == 'reference' else 'oracle port'} "
                                   f"({'OpenBLAS ' if m2['blas'] else 'non-BLAS '}par_loop/slice/all; {REF_BUILD}), asymmetric sweep member "
                                   f"n=(16,1024,4,16384) fp32 first-order, q=1..4, {m2['steps_done']} steps after 1 warm-up")}
        except Exception as exc:
            cpu_asym = {"value": None, "unit": "GB/s", "cores": 0, "kind": "port", "sample": f"failed: {exc}"}

    if e2e:
        e2e.pop("_a_host", None)
        if world > 1:
            e2e["host_bound_to_gpu_socket"] = bool(numa_bound)

    # ---- everything below is outside every timed region above: free the 16 GiB workload first ---------------------
    del a, cs, live
    torch.cuda.empty_cache()
    mg_check = None
    if world > 1 and not args.no_selfcheck:
        mg_check = multi_gpu_selfcheck(torch, dist, rank, world, dev)
    cfg5 = None
    if not args.no_cfg5:
        cfg5 = cfg5_leg(torch, dist, ttv_b200, rank, world, dev, args)
    sweep = None
    if world == 1 and not args.no_sweep:
        with ClockSampler(local) as sweep_clocks:
            sweep = sweep_leg(torch, ttv_b200, args)
        sweep["clocks"] = sweep_clocks.summary()
    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
                "config": {"workload": workload_name(world, args.nccl_reduce), "layout": "first-order", "per_gpu_tensor_bytes": sh.a_count * ELEM,
                           "timed_on": f"{world} x B200", "sampled_shape": list(sh.na_local), "sampled_bytes_per_step": total_bytes // world,
                           "sample_is": "the whole per-GPU 256^4 tensor",
                           "l2": "inputs (16 GiB per GPU) are larger than L2; no flush needed",
                           "timing": "CUDA events on the launching stream, max over ranks"},
                "gflops": round(total_flops / (ms_per_step * 1e-3) / 1e9, 1),
                "frac_of_measured_peak": round(value / (peak * world), 4),
                "frac_of_nominal_8000": round(value / (8000.0 * world), 4),
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clk}
        if cpu_asym is not None:
            line["cpu_baseline_asym"] = cpu_asym
        if sweep is not None:
            line["configs"] = sweep
        if cfg5 is not None:
            line["cfg5"] = cfg5
        if mg_check is not None:
            line["multi_gpu_selfcheck"] = mg_check
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def sweep_leg(torch, ttv_b200, args):
    """EVERY named BASELINE config, full size, device-resident: timed (CUDA events around each launch, A rotated over several
    buffers when it is not much larger than L2) and checked (sampled outputs against a host long-double dot on regenerated
    fibers; int32 bit-exact).  Not part of any timed region of the headline numbers."""
    from ttv_b200.measure import Arena, kernel_label, measure_config
    from ttv_b200.workloads import configs
    arena_a, arena_c = Arena(int(17.5e9)), Arena(int(8.8e9))
    rng = np.random.default_rng(20261017)
    rows, failures, checked, errors = [], 0, 0, []
    t0 = time.perf_counter()
    # the GPU has been idle while the CPU baseline ran: bring clocks and allocator back to working state before the first
    # config is timed (its kernels are 80 us long)
    measure_config("f32", [256, 256, 256, 64], [1, 2, 3, 4], 2, reps=20, warmup=5, check=False, arena_a=arena_a, arena_c=arena_c)
    for name, dt, na, pia, q, *rest in configs(args.sweep_set):
        try:
            pl = ttv_b200.plan(q, na, pia, dtype=dt)
            r = measure_config(dt, na, pia, q, reps=args.sweep_reps, warmup=2, arena_a=arena_a, arena_c=arena_c, rng=rng)
        except Exception as exc:
            errors.append(f"{name} {dt} q={q}: {exc}")
            continue
        failures += r.get("failures", 0)
        checked += r.get("checked", 0)
        rows.append({"name": name, "dtype": dt, "q": q, "view": [pl["outer"], pl["nq"], pl["inner"]], "kernel": kernel_label(pl),
                     "ms": round(r["ms_med"], 4), "gbs": round(r["gbs_med"], 1), "failures": r.get("failures", 0)})
    del arena_a, arena_c
    torch.cuda.empty_cache()
    if not rows:
        return {"n": 0, "errors": errors}
    gbs = sorted(x["gbs"] for x in rows)
    worst = min(rows, key=lambda x: x["gbs"])
    below = [f"{x['name']} {x['dtype']} q={x['q']}: {x['gbs']:.0f}" for x in rows if x["gbs"] < 0.8 * 8000.0]
    out = {"set": args.sweep_set, "n": len(rows), "min_gbs": worst["gbs"], "min_name": f"{worst['name']} {worst['dtype']} q={worst['q']}",
           "median_gbs": gbs[len(gbs) // 2], "geomean_gbs": round(float(np.exp(np.mean(np.log(gbs)))), 1),
           "below_0p8_nominal": below, "n_at_or_above_0p8_nominal": len(rows) - len(below),
           "parity_failures": failures, "samples_checked": checked, "errors": errors,
           "tolerance": "int32 bit-exact; float/complex |c - ref| <= 2 n_q eps sum|a_k||b_k| per component against a host long-double dot",
           "timing": (f"median of {args.sweep_reps} samples; a sample = ~2 ms of back-to-back launches inside one CUDA event pair, enqueued "
                      "behind a blocker kernel so that the host stays ahead of the GPU; A rotated over 2-4 buffers below 8 x L2"),
           "seconds": round(time.perf_counter() - t0, 1)}
    c12 = [x for x in rows if x["name"] == "cfg1" and x["q"] == 2]
    if c12:
        out["cfg1_q2"] = {"workload": "BASELINE configs[0]: 512^3 fp32, q=2, first-order", "ms": c12[0]["ms"], "gbs": c12[0]["gbs"],
                          "frac_of_nominal_8000": round(c12[0]["gbs"] / 8000.0, 4), "kernel": c12[0]["kernel"]}
    out["table"] = [[x["name"], x["dtype"], x["q"], x["gbs"], x["kernel"]] for x in rows]
    return out


def cfg5_leg(torch, dist, ttv_b200, rank, world, dev, args):
    """BASELINE configs[4]: order-3 fp64 n = (2048, 2048, 2048) (68.7 GB), STRONG scaling over the N ranks: sharded along mode 3,
    q = 1 and q = 2 are free splits (no communication), q = 3 contracts the split mode (n_q split; N > 1: exchange fused into
    the kernel's stores, N = 1: the plain kernel).  Launches are enqueued back to back (asynchronous), timed with CUDA events
    per rank; the line carries the max over ranks and the spread."""
    from ttv_b200 import selfcheck
    from ttv_b200.sharded import PeerExchange, make_shard, ttv_sharded
    na, pia, dt = [2048, 2048, 2048], [1, 2, 3], "f64"
    free_b, _ = torch.cuda.mem_get_info(dev)
    need = 8 * (2048 ** 3) // world + (2 << 30)
    if free_b < need:
        return {"skipped": f"needs {need >> 30} GiB of free HBM per GPU, {free_b >> 30} GiB free"}
    shards = {q: make_shard(q, na, pia, rank, world) for q in (1, 2, 3)}
    sh = shards[1]
    a = torch.empty(sh.a_count, dtype=torch.float64, device=dev)
    ttv_b200.fill(a, SEED_A, first=sh.a_offset)
    bs = {}
    for q in (1, 2, 3):
        bs[q] = torch.empty(na[q - 1], dtype=torch.float64, device=dev)
        ttv_b200.fill(bs[q], SEED_B + q)
    cs = {q: torch.full((shards[q].c_count,), float("nan"), dtype=torch.float64, device=dev) for q in (1, 2, 3)}
    exchange = None
    if world > 1 and not args.nccl_reduce:
        try:
            exchange = PeerExchange(shards[3].c_count, torch.float64, dev)
        except Exception as exc:
            print(f"bench.py: rank {rank}: cfg5: peer-memory exchange unavailable ({exc}); using ncclReduce", file=sys.stderr)
        ok = torch.tensor([1 if exchange is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            exchange = None
    out = {"workload": f"BASELINE configs[4]: order-3 fp64 n=(2048,2048,2048) first-order, strong scaling over {world} GPU(s), sharded along mode 3",
           "scaling": "strong", "per_gpu_tensor_bytes": sh.a_count * 8,
           "exchange": None if world == 1 else ("one kernel per GPU: stores over NVLink peer memory, in-kernel flag barrier, slot sum" if exchange is not None else "ncclReduce")}
    reps = max(5, min(args.steps, 20))
    rng = np.random.default_rng(55 + rank)
    for q in (1, 2, 3):
        def once():
            return ttv_sharded(q, a, na, pia, bs[q], rank=rank, world=world, c_local=cs[q], reduce_to=0, exchange=exchange, asynchronous=True)
        for _ in range(3):
            c, s = once()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            c, s = once()
        e1.record()
        torch.cuda.synchronize()
        ms_local = e0.elapsed_time(e1) / reps
        ms_all = [ms_local]
        if world > 1:
            t = torch.zeros(world, device=dev, dtype=torch.float64)
            t[rank] = ms_local
            dist.all_reduce(t)
            ms_all = [float(x) for x in t.tolist()]
        ms = max(ms_all)
        # parity of this rank's piece of C on sampled fibers (global indices: the slab was filled with first = its offset)
        bad = 0
        nchk = 0
        if not (s.kind == "nq" and world > 1 and rank != 0):
            nchk, bad, _ = selfcheck.check_product(c, dt, na, pia, q, SEED_A, bs[q], samples=32, rng=rng, c_first=s.c_offset)
        if world > 1:
            t = torch.tensor([bad, nchk], device=dev, dtype=torch.int64)
            dist.all_reduce(t)
            bad, nchk = int(t[0].item()), int(t[1].item())
        byt = 8 * (2048 ** 3 + 2048 + 2048 ** 2)
        pl = ttv_b200.plan(q, list(shards[q].na_local), pia, dtype=dt) if shards[q].count else {}
        from ttv_b200.measure import kernel_label
        forms = None
        if q == 3 and world > 1 and exchange is not None:
            # the same product with the two other exchange forms, timed the same way (none of them is the headline)
            forms = {}
            others = {"plain kernel + ncclReduce": None}
            try:
                others["scatter kernel + symmetric-memory barrier + reduce kernel"] = PeerExchange(shards[3].c_count, torch.float64, dev, single_kernel=False)
            except Exception:
                pass
            for fname, ex2 in others.items():
                for _ in range(3):
                    ttv_sharded(q, a, na, pia, bs[q], rank=rank, world=world, c_local=cs[q], reduce_to=0, exchange=ex2, asynchronous=True)
                dist.barrier()
                torch.cuda.synchronize()
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0.record()
                for _ in range(reps):
                    ttv_sharded(q, a, na, pia, bs[q], rank=rank, world=world, c_local=cs[q], reduce_to=0, exchange=ex2, asynchronous=True)
                f1.record()
                torch.cuda.synchronize()
                t = torch.tensor([f0.elapsed_time(f1) / reps], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                forms[fname] = round(float(t.item()), 4)
            forms["one kernel: scatter over NVLink + in-kernel barrier + slot sum"] = round(ms, 4)
            if exchange.timed_out():
                raise SystemExit(f"bench.py: rank {rank}: the single-kernel exchange timed out")
        out[f"q{q}"] = {"split": "free (no communication)" if shards[q].kind == "free" else "n_q split",
                        "ms": round(ms, 4), "gbs": round(byt / ms / 1e6, 1), "gbs_per_gpu": round(byt / ms / 1e6 / world, 1),
                        "frac_of_nominal_8000_per_gpu": round(byt / ms / 1e6 / world / 8000.0, 4),
                        "rank_ms_min": round(min(ms_all), 4), "rank_ms_max": round(max(ms_all), 4),
                        "kernel": (("ttv_col_exchange_kernel (product + scatter over NVLink + in-kernel barrier + slot sum, one launch)"
                                    if exchange.single_kernel else "ttv_col_scatter_kernel + symmetric-memory barrier + ttv_reduce_kernel")
                                   if (s.kind == "nq-scattered") else kernel_label(pl)),
                        "samples_checked": nchk, "parity_failures": bad}
        if forms:
            out[f"q{q}"]["exchange_forms_ms"] = forms
    del a, cs
    torch.cuda.empty_cache()
    return out


def multi_gpu_selfcheck(torch, dist, rank, world, dev):
    """Small integer-valued tensors through every exchange form of the sharded path -- free split, n_q split + NCCL all-reduce,
    n_q split fused with the exchange over peer memory -- compared bit for bit with numpy's tensordot of the global problem
    (int32, and float64 / complex64 holding small integers: every partial sum is exact).  Uneven splits included."""
    from ttv_b200 import selfcheck
    from ttv_b200.sharded import PeerExchange, make_shard, ttv_sharded
    cases = [((37, 21, 40), (1, 2, 3)), ((16, 33, 9, 12), (2, 1, 4, 3)), ((300, 17), (1, 2)), ((17, 300), (2, 1)), ((5, 6, 7, 19), (4, 3, 2, 1))]
    products, bad, fused = 0, 0, 0
    for npdt in (np.int32, np.float64, np.complex64):
        tdt = torch.from_numpy(np.zeros(1, npdt)).dtype
        try:
            ex = PeerExchange(40000, tdt, dev)
        except Exception:
            ex = None
        rng = np.random.default_rng(99)                     # the same data on every rank
        for na, pia in cases:
            n = int(np.prod(na))
            a_full = rng.integers(-6, 7, n).astype(npdt)
            for q in range(1, len(na) + 1):
                b = rng.integers(-6, 7, na[q - 1]).astype(npdt)
                want = selfcheck.full_product(a_full, na, pia, q, b)
                sh = make_shard(q, na, pia, rank, world)
                a_loc = torch.from_numpy(a_full[sh.a_offset: sh.a_offset + sh.a_count].copy()).to(dev)
                tb = torch.from_numpy(b).to(dev)
                c, s2 = ttv_sharded(q, a_loc, na, pia, tb, rank=rank, world=world, reduce_to=None)
                got = c.cpu().numpy()
                ref = want[s2.c_offset: s2.c_offset + s2.c_count] if s2.kind == "free" else want
                bad += 0 if np.array_equal(got, ref) else 1
                products += 1
                if s2.kind != "free" and ex is not None and na[s2.mode - 1] >= world:
                    for _ in range(2):                      # both halves of the workspace
                        c2, s3 = ttv_sharded(q, a_loc, na, pia, tb, rank=rank, world=world, exchange=ex)
                        bad += 0 if np.array_equal(c2.cpu().numpy(), want[s3.c_offset: s3.c_offset + s3.c_count]) else 1
                        products += 1
                        fused += 1
    t = torch.tensor([products, bad, fused], device=dev, dtype=torch.int64)
    dist.all_reduce(t)
    return {"products_checked_all_ranks": int(t[0].item()), "failures": int(t[1].item()), "fused_exchange_products": int(t[2].item()),
            "against": "numpy tensordot of the global tensor, bit for bit (integer-valued data)"}


def local_product(ttv_b200, q, a, shard, pia, b, c):
    """this rank's kernel launch for product q (no collective): what the roofline of the kernel is measured on"""
    na = list(shard.na_local)
    bq = b[shard.begin: shard.begin + shard.count] if shard.kind == "nq" else b
    nc = ttv_b200.generate_output_shape(na, q); pic = ttv_b200.generate_output_layout(pia, q)
    ttv_b200.ttv_lowlevel(q, len(na), a, na, ttv_b200.generate_strides(na, pia), pia, bq, [int(bq.shape[0])], c, nc,
                          ttv_b200.generate_strides(nc, pic), pic, flags=2)


def verify_sample(torch, cs, shards, na_global, bs, rank, world):
    """sampled outputs of every product against a host long-double dot on regenerated fibers (ttv_b200/selfcheck.py: the
    fill generator restated in numpy -- no code under oracle/ runs on this arm outside the cpu_baseline leg)"""
    from ttv_b200 import selfcheck
    rng = np.random.default_rng(1234 + rank)
    for q in range(1, ORDER + 1):
        sh = shards[q]
        if (sh.kind == "nq" and (world > 1 and rank != 0)) or sh.c_count == 0:
            continue
        n, bad, worst = selfcheck.check_product(cs[q], DTYPE, na_global, [1, 2, 3, 4], q, SEED_A, bs[q], samples=64, rng=rng, c_first=sh.c_offset)
        if bad:
            raise SystemExit(f"bench.py: parity check failed for q={q} on rank {rank}: {bad} of {n} sampled outputs outside the tolerance "
                             f"(worst |err| / tol = {worst:.3g})")


def measure_e2e(torch, dist, ttv_b200, args, a, bs, cs, shards, na_global, pia, rank, world, dev, total_bytes, exchange=None):
    """Same step, inputs in pinned HOST memory; every step copies its inputs (this rank's slab of A and the four
    vectors) to the device and reads the four results back inside the timed region.  N = 1: four drop-in calls of the
    low-level interface on a tensor that keeps its device copy between products (invalidated every step: A crosses PCIe
    once per step, streamed under the first product's kernels).  Beside it: `multi_value` (ONE call of ttv_b200_multi for the
    four products) and `per_call_value` (four independent calls on a plain host tensor, A crossing PCIe four times).
    N > 1 stages explicitly because the n_q-split reduce runs on device buffers."""
    sh = shards[1]
    steps = max(1, args.e2e_steps)
    a_host = torch.empty(sh.a_count, dtype=torch.float32, pin_memory=True)
    a_host.copy_(a)                                      # same synthetic data as the device run
    b_host = {q: bs[q].cpu().pin_memory() for q in bs}
    c_host = {q: torch.empty(shards[q].c_count, dtype=torch.float32, pin_memory=True) for q in cs}
    h2d = sh.a_count * ELEM + sum((shards[q].count if shards[q].kind == "nq" else na_global[q - 1]) * ELEM for q in range(1, ORDER + 1))
    if exchange is not None:      # fused exchange: every rank reads back its block of the n_q-split product
        d2h = sum((shards[q].c_count // world if shards[q].kind == "nq" else shards[q].c_count) * ELEM for q in range(1, ORDER + 1))
    else:
        d2h = sum(shards[q].c_count * ELEM for q in range(1, ORDER + 1) if not (shards[q].kind == "nq" and rank != 0))

    qs = list(range(1, ORDER + 1))
    a_np = a_host.numpy()
    b_np = [b_host[q].numpy() for q in qs]
    c_np = [c_host[q].numpy() for q in qs]

    resident = ttv_b200.Resident() if world == 1 else None

    def e2e_step():
        if world == 1:
            # The reference-facing call, product by product: four drop-in calls of the low-level interface with HOST pointers,
            # on a tensor that keeps its copy in HBM (ttv_b200_run_resident -- what `A(q) * b` does for a tlib::ttv::tensor
            # after A.keep_on_device()).  New data arrives with every step (invalidate), so the first product streams all of
            # A across PCIe under its own kernels; the other three read HBM and move only b and C.
            resident.invalidate()
            for q in qs:
                na = list(sh.na_local)
                nc = ttv_b200.generate_output_shape(na, q); pic = ttv_b200.generate_output_layout(pia, q)
                resident.ttv_lowlevel(q, ORDER, a_np, na, ttv_b200.generate_strides(na, pia), pia, b_np[q - 1], [na[q - 1]],
                                      c_np[q - 1], nc, ttv_b200.generate_strides(nc, pic), pic)
        else:
            from ttv_b200.sharded import ttv_sharded
            a.copy_(a_host, non_blocking=True)                      # this rank's slab, once per step
            for q in qs:
                bs[q].copy_(b_host[q], non_blocking=True)
                c, s = ttv_sharded(q, a, na_global, pia, bs[q], rank=rank, world=world, c_local=cs[q], reduce_to=0, exchange=exchange)
                if s.kind == "nq-scattered":
                    c_host[q][: s.c_count].copy_(c, non_blocking=True)          # this rank's block of C
                elif not (s.kind == "nq" and rank != 0):
                    c_host[q].copy_(cs[q], non_blocking=True)
            torch.cuda.synchronize()

    def e2e_step_multi():
        """ONE C-ABI call for the four products (ttv_b200_multi, an entry the reference has no counterpart for)"""
        ttv_b200.ttv_multi(qs, a_np, list(sh.na_local), pia, b_np, c_np)

    def e2e_step_per_call():
        """the strict per-product variant: every product is its own C-ABI call with host buffers (A crosses PCIe 4x)"""
        for q in qs:
            na = list(shards[q].na_local)
            nc = ttv_b200.generate_output_shape(na, q); pic = ttv_b200.generate_output_layout(pia, q)
            ttv_b200.ttv_lowlevel(q, ORDER, a_np, na, ttv_b200.generate_strides(na, pia), pia, b_np[q - 1], [na[q - 1]],
                                  c_np[q - 1], nc, ttv_b200.generate_strides(nc, pic), pic)

    e2e_step()                                            # warm-up (also sizes the staging buffers)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        hb = torch.tensor([h2d, d2h], device=dev, dtype=torch.int64)
        dist.all_reduce(hb)
        h2d, d2h = int(hb[0].item()), int(hb[1].item())
    # the host-path results must be the device-resident ones up to summation order (the host path streams A in chunks, so
    # its tiles differ): |diff| <= 2 n_q (eps/2) sum_k |a_k||b_k| <= n_q eps sum_k |b_k| because |a| < 1
    if world == 1:
        for q in qs:
            nq = na_global[q - 1]
            tol = nq * float(np.finfo(np.float32).eps) * float(b_host[q].abs().sum())
            err = float((c_host[q] - cs[q].cpu()).abs().max())
            if not err <= tol:
                raise SystemExit(f"bench.py: e2e result of q={q} differs from the device-resident result: {err} > {tol}")
    out = {"value": round(total_bytes / (dt / steps) / 1e9, 2), "unit": "GB/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": round(dt / steps * 1e3, 2), "steps": steps,
           "path": ("four drop-in calls per step (q = 1..4) of the low-level interface with pinned HOST buffers on a tensor that keeps its "
                    "copy in HBM (ttv_b200_run_resident = tlib::ttv::tensor::keep_on_device behind A(q)*b), invalidated at the start "
                    "of every step: A crosses PCIe once per step" if world == 1
                    else "pinned host -> device copy of the slab once per step, four sharded products, device -> host"),
           "bound": "PCIe: the step moves 16 GiB of A per GPU over a Gen5 x16 link"}
    if world == 1:
        out["_a_host"] = a_np          # handed to the cpu_baseline leg (popped before the line is printed)
        resident.close()
        e2e_step_multi()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_step_multi()
        torch.cuda.synchronize()
        out["multi_value"] = round(total_bytes / (time.perf_counter() - t0) / 1e9, 2)
        out["multi_path"] = "ttv_b200_multi: one C-ABI call runs the four products on one upload of A"
        e2e_step_per_call()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_step_per_call()
        torch.cuda.synchronize()
        dt1 = time.perf_counter() - t0
        out["per_call_value"] = round(total_bytes / dt1 / 1e9, 2)
        out["per_call_path"] = "four independent ttv_b200_f32 calls on a plain host tensor: A crosses PCIe four times per step"
        out["per_call_h2d_bytes_per_step"] = int(4 * sh.a_count * ELEM + sum(na_global) * ELEM)
    return out


_JSON_FD = None


def _guard_stdout():
    """Libraries print to the process's stdout behind Python's back (NCCL's version banner at NCCL_DEBUG=VERSION/WARN,
    torchrun notices).  The contract is ONE JSON line on stdout: keep a private duplicate of fd 1 for that line and point
    fd 1 itself at stderr for everything else."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="N > 1: do not pin the rank to the CPUs next to its GPU")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the every-named-config leg (N = 1)")
    ap.add_argument("--sweep-set", default="named", help="config family of ttv_b200/workloads.py for the sweep leg")
    ap.add_argument("--sweep-reps", type=int, default=7)
    ap.add_argument("--no-cfg5", action="store_true", help="skip the 2048^3 fp64 strong-scaling leg")
    ap.add_argument("--no-selfcheck", action="store_true", help="N > 1: skip the small-tensor parity check of the sharded path")
    ap.add_argument("--nccl-reduce", action="store_true", help="N > 1: plain kernel + ncclReduce for the n_q-split product instead of the fused exchange")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_own_arm(args)


if __name__ == "__main__":
    main()
