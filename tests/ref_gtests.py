"""Compiles the reference's OWN test sources (test/src/*.cpp of bassoy/ttv, unmodified, from where they lie under
/root/reference) against THIS repo's include/tlib headers, libttv_b200.so and the GoogleTest shim in tests/gtest_shim.

    tests/_refbin/ref_gtests_host   gtest_tlib_{layout,shape,strides,workload}.cpp   (pure host logic, runs anywhere)
    tests/_refbin/ref_gtests_gpu    gtest_tlib_{ttv,mtv}.cpp                         (19 policy combinations, needs a GPU)
    tests/_refbin/ref_interface{1,2,3}   the reference's example/interface{1,2,3}.cpp, unmodified (needs a GPU)
    tests/_refbin/iface_check       tests/cpp/iface_check.cpp: the two high-level interfaces on arbitrary cases (needs a GPU)

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


def _build_plain(name, source):
    """one C++ source against include/ and libttv_b200.so"""
    os.makedirs(BIN, exist_ok=True)
    out = os.path.join(BIN, name)
    cmd = ["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), source,
           "-L" + os.path.join(ROOT, "ttv_b200"), "-lttv_b200", "-Wl,-rpath,$ORIGIN/../../ttv_b200", "-o", out]
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError("building " + source + " against the new headers failed:\n" + " ".join(cmd) + "\n" + r.stderr[-4000:])
    return out


def build_iface_check():
    """this repo's own driver of the high-level interfaces: needs no reference source"""
    return _build_plain("iface_check", os.path.join(ROOT, "tests", "cpp", "iface_check.cpp"))


def build_all():
    built = [build_iface_check()]
    if not reference_present():
        return built
    built += [_build("ref_gtests_host", HOST_SOURCES), _build("ref_gtests_gpu", GPU_SOURCES)]
    # the reference's three example programs, compiled UNMODIFIED from where they lie
    built += [_build_plain(f"ref_interface{i}", os.path.join(REF, "example", f"interface{i}.cpp")) for i in (1, 2, 3)]
    return built


def binary(name):
    path = os.path.join(BIN, name)
    return path if os.path.exists(path) else None
