"""Builds ttv_b200/libttv_b200.so (the C-ABI shared library) with nvcc for sm_100a, in-tree.

    python -m ttv_b200.build            # build if sources are newer than the library
    python -m ttv_b200.build --force
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libttv_b200.so")
SOURCES = ["api.cu", "launch.cu", "plan.cpp"]
HEADERS = ["plan.h", "launch.h", "kernels.cuh", "numeric.cuh", os.path.join("..", "..", "include", "ttv_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall",
    "-shared", "--cudart", "static",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise FileNotFoundError("nvcc not found")


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, f) for f in SOURCES] + ["-o", LIB]
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)       # the image's CC wrapper lacks pieces nvcc's host pass needs
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
