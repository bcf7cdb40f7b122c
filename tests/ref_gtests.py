"""Compiles the reference's OWN test sources (test/src/*.cpp of bassoy/ttv, unmodified, from where they lie under
/root/reference) against THIS repo's include/tlib headers, libttv_b200.so and the GoogleTest shim in tests/gtest_shim.

    tests/_refbin/ref_gtests_host   gtest_tlib_{layout,shape,strides,workload}.cpp   (pure host logic, runs anywhere)
    tests/_refbin/ref_gtests_gpu    gtest_tlib_{ttv,mtv}.cpp                         (19 policy combinations, needs a GPU)

The binaries are built in the container (build()) and travel to the GPU box with the snapshot (tests/_refbin is
git-ignored, not gpurun-ignored); no reference source is copied into the repo.
"""
from __future__ import annotations

import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
BIN = os.path.join(ROOT, "tests", "_refbin")
HOST_SOURCES = ["gtest_tlib_layout.cpp", "gtest_tlib_shape.cpp", "gtest_tlib_strides.cpp", "gtest_tlib_workload.cpp"]
GPU_SOURCES = ["gtest_tlib_ttv.cpp", "gtest_tlib_mtv.cpp"]


def reference_present() -> bool:
    return os.path.isdir(os.path.join(REF, "test", "src"))


def _build(name, sources):
    os.makedirs(BIN, exist_ok=True)
    out = os.path.join(BIN, name)
    libdir = os.path.join(ROOT, "ttv_b200")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Wextra",
           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "tests", "gtest_shim"),
           "-I" + os.path.join(REF, "test", "include")] + \
          [os.path.join(REF, "test", "src", s) for s in sources + ["main.cpp"]] + \
          ["-L" + libdir, "-lttv_b200", "-Wl,-rpath,$ORIGIN/../../ttv_b200", "-o", out]
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError("building the reference's tests against the new headers failed:\n" + " ".join(cmd) + "\n" + r.stderr[-4000:])
    return out


def build_all():
    if not reference_present():
        return []
    return [_build("ref_gtests_host", HOST_SOURCES), _build("ref_gtests_gpu", GPU_SOURCES)]


def binary(name):
    path = os.path.join(BIN, name)
    return path if os.path.exists(path) else None
