"""ttv_b200/csrc/hostcopy.cpp -- the memcpy the copy threads of the host path run (streaming stores) -- compiled on its own
and checked for every kind of ragged size and alignment (the GPU tests only see whole chunks from page-aligned buffers)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "ttv_b200", "csrc", "hostcopy.cpp")
OUT = os.path.join(ROOT, "tests", "_refbin", "libhostcopy_test.so")


@pytest.fixture(scope="module")
def host_copy():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", SRC, "-o", OUT], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    lib = C.CDLL(OUT)
    fn = getattr(lib, "_ZN4ttvb9host_copyEPvPKvm")           # ttvb::host_copy(void*, void const*, unsigned long)
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    return fn


def test_host_copy_ragged_sizes_and_alignments(host_copy):
    rng = np.random.default_rng(3)
    src_buf = rng.integers(0, 256, (1 << 21) + 4096, dtype=np.uint8)
    sizes = [0, 1, 31, 32, 33, 127, 128, 4095, 16383, 16384, 16385, 16384 + 31, 65536 + 97, 1 << 20, (1 << 20) + 1, (1 << 21) - 13]
    for n in sizes:
        for so in (0, 1, 17, 32, 63):
            for do in (0, 1, 15, 31, 32, 33):
                dst_buf = np.full(n + 256, 0xAB, np.uint8)
                host_copy(dst_buf.ctypes.data + 64 + do, src_buf.ctypes.data + so, n)
                assert np.array_equal(dst_buf[64 + do: 64 + do + n], src_buf[so: so + n]), (n, so, do)
                assert np.all(dst_buf[: 64 + do] == 0xAB) and np.all(dst_buf[64 + do + n:] == 0xAB), ("wrote outside", n, so, do)


@pytest.mark.parametrize("tsan", [False, True])
def test_copy_pool_with_concurrent_callers(tsan):
    """ttv_b200/csrc/copy_pool.h under several concurrent callers (tests/copy_pool_harness.cpp); the second build runs
    under ThreadSanitizer, which must report no data race"""
    exe = os.path.join(ROOT, "tests", "_refbin", "copy_pool_harness" + ("_tsan" if tsan else ""))
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    cmd = ["g++", "-O1" if tsan else "-O2", "-g", "-std=c++17", "-pthread", os.path.join(ROOT, "tests", "copy_pool_harness.cpp"), SRC, "-o", exe]
    if tsan:
        cmd.insert(1, "-fsanitize=thread")
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if tsan and r.returncode != 0 and "tsan" in (r.stderr + r.stdout).lower():
        pytest.skip("ThreadSanitizer runtime is not installed")
    assert r.returncode == 0, r.stderr[-3000:]
    args = ["3", "5", "12"] if tsan else ["4", "7", "40"]
    r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=600)
    if tsan and "FATAL: ThreadSanitizer" in r.stderr:      # e.g. an address-space layout TSan cannot map in this container
        pytest.skip("ThreadSanitizer cannot run here: " + r.stderr.splitlines()[0])
    assert r.returncode == 0 and " 0 bad" in r.stdout, r.stdout + r.stderr[-3000:]
    assert "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[-4000:]
