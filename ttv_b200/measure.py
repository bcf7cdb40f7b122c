"""Device-resident timing + sampled parity of ONE product, shared by bench.py's sweep leg, tools/sweep.py and the
full-size GPU tests.  CUDA events on the launching stream around every launch; inputs larger than L2 or rotated over
several buffers; results checked against ttv_b200.selfcheck (host long-double dot on regenerated fibers).
torch is used for device memory, streams and events only."""
from __future__ import annotations

import numpy as np

from . import api, selfcheck
from .workloads import SEED_A, SEED_B, SIZE, algo_bytes

L2_BYTES = 126 * 2 ** 20
KERNEL_NAMES = {0: "auto", 1: "dot", 2: "col", 3: "stream", 4: "colx", 5: "dotf", 6: "strided", 7: "colt", 8: "streamk", 9: "dotp", 10: "colf"}


def _torch_dtype(dt):
    import torch
    return {"f32": torch.float32, "f64": torch.float64, "c64": torch.complex64, "c128": torch.complex128,
            "i32": torch.int32, "i64": torch.int64}[dt]


class Arena:
    """one device allocation carved into typed views (the sweep visits ~200 shapes of up to 17 GB: no allocator churn)"""

    def __init__(self, nbytes: int, device="cuda"):
        import torch
        self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        self.off = 0

    def reset(self):
        self.off = 0

    def take(self, count: int, dt: str):
        nbytes = int(count) * SIZE[dt]
        start = (self.off + 255) // 256 * 256
        if start + nbytes > self.buf.numel():
            raise MemoryError(f"arena of {self.buf.numel()} bytes cannot hold {nbytes} more")
        self.off = start + nbytes
        return self.buf[start: start + nbytes].view(_torch_dtype(dt))


_BLOCKER = []


def _blocker():
    """1 GiB of scratch whose refill keeps the GPU busy for a few hundred microseconds"""
    if not _BLOCKER:
        import torch
        _BLOCKER.append(torch.empty(1 << 28, dtype=torch.float32, device="cuda"))
    return _BLOCKER[0]


def kernel_label(pl: dict) -> str:
    """a readable name of what the chooser picked (plan dict of ttv_b200.plan)"""
    fam = KERNEL_NAMES.get(pl.get("kernel"), "?")
    bits = [f"ttv_{fam}_kernel", f"vec{pl.get('vec')}", f"tile({pl.get('to')},{pl.get('ty')},{pl.get('tx')})",
            f"batch({pl.get('nu')},{pl.get('ku')})"]
    if pl.get("ksplit", 1) > 1:
        bits.append(f"ksplit{pl['ksplit']}+ttv_reduce_kernel")
    return " ".join(bits)


def measure_config(dt: str, na, pia, q: int, *, wa=None, reps: int = 10, warmup: int = 3, check: bool = True,
                   samples: int = 64, arena_a: Arena | None = None, arena_c: Arena | None = None, rng=None, **opts) -> dict:
    """Times C = A x_q b on synthetic device-resident data and (check=True) verifies sampled outputs of the LAST launch.
    Returns ms_med / ms_min / gbs_med / gbs_best / bytes / checked / failures / worst_err_over_tol."""
    import torch
    n = int(np.prod(na, dtype=object))
    s = SIZE[dt]
    nq = int(na[q - 1])
    span = n if wa is None else 1 + sum((int(e) - 1) * int(w) for e, w in zip(na, wa))
    copies = max(1, min(4, -(-8 * L2_BYTES // (span * s))))           # rotate when A is not much larger than L2
    tdt = _torch_dtype(dt)
    if arena_a is not None:
        arena_a.reset()
    if arena_c is not None:
        arena_c.reset()
    As = []
    for i in range(copies):
        a = arena_a.take(span, dt) if arena_a is not None else torch.empty(span, dtype=tdt, device="cuda")
        api.fill(a, SEED_A + i)
        As.append(a)
    b = torch.empty(nq, dtype=tdt, device="cuda")
    api.fill(b, SEED_B)
    nc = api.generate_output_shape(na, q); pic = api.generate_output_layout(pia, q)
    flags = 2 if wa is None else 2 | 8
    wa_ = api.generate_strides(na, pia) if wa is None else list(wa)
    wc = api.generate_strides(nc, pic)
    n_out = n // nq
    c = arena_c.take(n_out, dt) if arena_c is not None else torch.empty(n_out, dtype=tdt, device="cuda")
    # poison C: a kernel that skips outputs cannot pass the check
    if tdt.is_floating_point or tdt.is_complex:
        c.fill_(float("nan"))
    else:
        c.fill_(0x7FFFFFFF)

    # one prepared call per buffer: a launch then costs a few microseconds of host time, so the host stays ahead of the GPU
    # even for the 80 us kernels of a 512 MiB tensor
    runs = [api.prepared_lowlevel(q, len(na), a, na, wa_, pia, b, [nq], c, nc, wc, pic, flags=flags, **opts) for a in As]
    for i in range(warmup):
        runs[i % copies]()
    torch.cuda.synchronize()
    byt = algo_bytes(dt, na, q)
    # Timing.  Events around a single launch measure the HOST when the kernel is shorter than the time the host needs to
    # enqueue it (the first event fires on an idle GPU and then waits for the launch).  So every sample is a train of
    # launches inside one event pair, enqueued behind a blocker kernel that keeps the GPU busy while the host runs ahead:
    # elapsed / train = time per product in a stream of such products, launch gaps included.
    est_ms = max(byt / 7.0e9, 0.004)
    train = int(max(1, min(64, round(2.0 / est_ms))))                 # ~2 ms of kernels per sample
    samples_ms = []
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if est_ms < 1.0:
            api.fill(_blocker(), 1)          # ~0.4 ms of GPU work in front of the event pair: the host gets its head start
        e0.record()
        for i in range(train):
            runs[(r * train + i) % copies]()
        e1.record()
        samples_ms.append((e0, e1))
    torch.cuda.synchronize()
    ts = sorted(e0.elapsed_time(e1) / train for e0, e1 in samples_ms)
    last = (reps * train - 1) % copies
    out = {"ms_med": ts[len(ts) // 2], "ms_min": ts[0], "gbs_med": byt / ts[len(ts) // 2] / 1e6, "gbs_best": byt / ts[0] / 1e6,
           "bytes": byt, "copies": copies, "train": train}
    if check and wa is None:
        checked, bad, worst = selfcheck.check_product(c, dt, na, pia, q, SEED_A + last, b, samples=samples, rng=rng)
        out.update(checked=checked, failures=bad, worst_err_over_tol=worst)
    return out
