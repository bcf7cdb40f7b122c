"""Builds ttv_b200/libttv_b200.so (the C-ABI shared library) with nvcc for sm_100a, in-tree.

    python -m ttv_b200.build            # build if sources are newer than the library
    python -m ttv_b200.build --force [-v]

launch.cu is compiled once per element type (-DTTVB_DTYPE=k) plus once for its dtype-independent part; the
translation units are compiled in parallel and linked into one shared library with a static CUDA runtime.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(PKG, "libttv_b200.so")
HEADERS = ["plan.h", "launch.h", "copy_pool.h", "kernels.cuh", "stream_kernel.cuh", "colx_kernel.cuh", "colr_kernel.cuh", "dotf_kernel.cuh", "strided_kernel.cuh", "numeric.cuh", os.path.join("..", "..", "include", "ttv_b200.h")]
N_DTYPES = 6

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ARCH + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-O3,-Wall"]


def units():
    """(source, object, extra flags)"""
    out = [("api.cu", "api.o", []), ("plan.cpp", "plan.o", []), ("hostcopy.cpp", "hostcopy.o", []), ("launch.cu", "launch.o", [])]
    out += [("launch.cu", f"launch_dtype{k}.o", [f"-DTTVB_DTYPE={k}"]) for k in range(N_DTYPES)]
    return out


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return "nvcc"


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in ["api.cu", "plan.cpp", "hostcopy.cpp", "launch.cu"] + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def _env():
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)       # the image's CC wrapper lacks pieces nvcc's host pass needs
    return env


def _compile(unit, verbose):
    src, obj, extra = unit
    cmd = [nvcc()] + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", os.path.join(OBJ, obj)]
    r = subprocess.run(cmd, env=_env(), capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return r.stderr


def build(force: bool = False, verbose: bool = False, defines=(), out: str = LIB) -> str:
    """defines/out: build an experimental variant beside the product library (e.g. defines=["-DTTVB_MIN_CTAS=4"])"""
    if not force and out == LIB and not stale():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    todo = [(src, obj, list(extra) + list(defines)) for src, obj, extra in units()]
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as pool:
        logs = list(pool.map(lambda u: _compile(u, verbose), todo))
    if verbose:
        sys.stderr.write("".join(logs))
    cmd = [nvcc()] + ARCH + ["-shared", "--cudart", "static"] + [os.path.join(OBJ, u[1]) for u in units()] + ["-o", out]
    r = subprocess.run(cmd, env=_env(), capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
