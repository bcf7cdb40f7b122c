"""Sampled parity check for tensors that are too large for a host-side oracle run ("too-big-for-host configs").

The synthetic tensors of the benchmarks are written by ttv_b200_fill, a counter-based generator: element j of a buffer
filled with `seed` is a pure function of (seed, j) (csrc/numeric.cuh splitmix64 / unit_pm1, csrc/kernels.cuh synth<T>).
This module restates that generator in numpy and uses it to recompute SAMPLED fibers of A on the host: an output C[o][i]
of the canonical view is  sum_k A[(o * n_q + k) * inner + i] * b[k],  evaluated here in long double (integers: exactly, with
wrap-around) and compared with what the device produced --

    integers          bit-exact
    float / complex   |c - c_ref| <= 2 * n_q * eps * sum_k |a_k| |b_k|   per component, eps = 2^-24 / 2^-53
                      (the tolerance BASELINE.json's north_star states)

It is the product's own self-check (bench.py, tools/sweep.py, the full-size GPU tests); it does not touch oracle/.
"""
from __future__ import annotations

import numpy as np

_EPS = {"f32": 2.0 ** -24, "c64": 2.0 ** -24, "f64": 2.0 ** -53, "c128": 2.0 ** -53}
NP_DTYPE = {"f32": np.float32, "f64": np.float64, "c64": np.complex64, "c128": np.complex128, "i32": np.int32, "i64": np.int64}


def _splitmix64(x):
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _unit_pm1(u):
    return (u >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0


def synth(dtype: str, seed: int, idx):
    """element `idx` (array of flat indices) of a buffer written by ttv_b200_fill(dtype, seed)"""
    j = np.asarray(idx, dtype=np.uint64)
    s = np.uint64(seed)
    if dtype == "f32":
        return _unit_pm1(_splitmix64(s ^ j)).astype(np.float32)
    if dtype == "f64":
        return _unit_pm1(_splitmix64(s ^ j))
    if dtype in ("c64", "c128"):
        with np.errstate(over="ignore"):
            re = _unit_pm1(_splitmix64(s ^ (np.uint64(2) * j)))
            im = _unit_pm1(_splitmix64(s ^ (np.uint64(2) * j + np.uint64(1))))
        if dtype == "c64":
            return (re.astype(np.float32) + 1j * im.astype(np.float32)).astype(np.complex64)
        return re + 1j * im
    if dtype in ("i32", "i64"):
        return ((_splitmix64(s ^ j) % np.uint64(17)).astype(np.int64) - 8).astype(NP_DTYPE[dtype])
    raise ValueError(dtype)


def view_of(na, pia, q):
    """(outer, n_q, inner) of the canonical view (csrc/plan.h)"""
    k = list(pia).index(q)
    inner = 1
    for m in pia[:k]:
        inner *= int(na[m - 1])
    outer = 1
    for m in pia[k + 1:]:
        outer *= int(na[m - 1])
    return outer, int(na[q - 1]), inner


def sample_indices(n_out: int, nq: int, samples: int, rng) -> np.ndarray:
    """flat output indices to check: the first and the last output plus random ones; fewer when the fibers are long (the
    host regenerates samples * n_q elements: about 2^25 at most, never fewer than two fibers)"""
    samples = int(max(2, min(samples, (1 << 25) // max(1, nq))))
    picks = {0, n_out - 1}
    if samples > 2:
        picks.add(n_out // 2)
        picks.update(int(x) for x in rng.integers(0, n_out, samples - 2))
    return np.array(sorted(picks), dtype=np.int64)


def expected(dtype: str, view, js, seed_a: int, b_host, c_first: int = 0):
    """(reference values, tolerances) of the outputs js (flat indices into THIS rank's C; c_first is the offset of the local
    C inside the global one and `view` the GLOBAL view: a sharded run fills its slab with first = its offset in the global A,
    so the elements of a fiber are regenerated from their global indices).
    b_host: the vector actually used (numpy).  Tolerance 0 for integers."""
    outer, nq, inner = view
    k = np.arange(nq, dtype=np.int64)
    want, tol = [], []
    integer = dtype in ("i32", "i64")
    # long double for ordinary fibers; very long ones (n_q > 2^16) are summed pairwise in double (numpy's sum), whose error
    # ~ log2(n_q) 2^-53 sum|a||b| is still far below the tolerance 2 n_q eps sum|a||b| -- x87 arithmetic on 10^8 elements
    # would take longer than the whole bench
    wide = np.longdouble if nq <= (1 << 16) else np.float64
    cwide = np.clongdouble if nq <= (1 << 16) else np.complex128
    if integer:
        bw = b_host.astype(np.int64)
    elif dtype in ("c64", "c128"):
        bw = b_host.astype(cwide)
    else:
        bw = b_host.astype(wide)
    CH = 1 << 20                                       # very long fibers are regenerated in cache-sized pieces
    for j in js:
        o, i = divmod(int(j) + c_first, inner)
        acc_i, acc_f, acc_abs = 0, 0.0, 0.0
        for k0 in range(0, nq, CH):
            kk = k[k0:k0 + CH]
            fiber = synth(dtype, seed_a, (o * nq + kk) * inner + i)
            bk = bw[k0:k0 + CH]
            if integer:
                with np.errstate(over="ignore"):
                    acc_i += int(np.sum(fiber.astype(np.int64) * bk))
            elif dtype in ("c64", "c128"):
                f = fiber.astype(cwide)
                acc_f = acc_f + np.sum(f * bk)
                acc_abs += float(np.sum(np.abs(f) * np.abs(bk)))
            else:
                f = fiber.astype(wide)
                acc_f = acc_f + np.sum(f * bk)
                acc_abs += float(np.sum(np.abs(f) * np.abs(bk)))
        if integer:
            bits = 32 if dtype == "i32" else 64
            acc_i &= (1 << bits) - 1
            if acc_i >= 1 << (bits - 1):
                acc_i -= 1 << bits
            want.append(acc_i); tol.append(0.0)
        else:
            want.append(complex(acc_f) if dtype in ("c64", "c128") else float(acc_f))
            tol.append(2.0 * nq * _EPS[dtype] * acc_abs + 1e-300)
    return want, tol


def check_product(c, dtype: str, na, pia, q: int, seed_a: int, b, *, samples: int = 64, rng=None,
                  c_first: int = 0, view=None):
    """Compares sampled elements of the device result `c` (flat torch tensor: this rank's C) with the host recomputation.
    Returns (number of samples, number of failures, worst |err| / tol).  Never raises on a mismatch: callers count."""
    import torch
    rng = rng if rng is not None else np.random.default_rng(1234)
    view = view if view is not None else view_of(na, pia, q)
    n_out = int(c.numel())
    if n_out == 0:
        return 0, 0, 0.0
    js = sample_indices(n_out, view[1], samples, rng)
    got = c[torch.from_numpy(js).to(c.device)].cpu().numpy()
    b_host = b.cpu().numpy() if hasattr(b, "cpu") else np.asarray(b)
    want, tol = expected(dtype, view, js, seed_a, b_host, c_first=c_first)
    bad, worst = 0, 0.0
    for g, w, t in zip(got, want, tol):
        if t == 0.0:
            ok = int(g) == int(w)
            ratio = 0.0 if ok else float("inf")
        elif dtype in ("c64", "c128"):
            err = max(abs(complex(g).real - w.real), abs(complex(g).imag - w.imag))
            ok = err <= t
            ratio = err / t
        else:
            err = abs(float(g) - w)
            ok = err <= t                      # (NaN fails)
            ratio = err / t if err == err else float("inf")
        bad += 0 if ok else 1
        worst = max(worst, ratio)
    return len(js), bad, worst


def full_product(a_flat, na, pia, q: int, b):
    """C = A x_q b of a SMALL host tensor with numpy, flat in the output layout (pia without q): the memory of a packed tensor
    with layout pia is a C-ordered array whose axes are the modes pia[p-1], ..., pia[0] (slowest first)."""
    p = len(na)
    modes_slow_to_fast = [int(m) for m in reversed(list(pia))]
    arr = np.asarray(a_flat).reshape([int(na[m - 1]) for m in modes_slow_to_fast])
    axis = modes_slow_to_fast.index(int(q))
    out = np.tensordot(arr, np.asarray(b), axes=([axis], [0]))
    return np.ascontiguousarray(out).reshape(-1).astype(np.asarray(a_flat).dtype, copy=False)
