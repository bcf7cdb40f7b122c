"""Builds ttv_b200/libttv_b200.so (the C-ABI shared library) with nvcc for sm_100a, in-tree.

    python -m ttv_b200.build            # build if sources are newer than the library
    python -m ttv_b200.build --force [-v]

launch.cu is compiled once per element type (-DTTVB_DTYPE=k) plus once for its dtype-independent part; the
translation units are compiled in parallel and linked into one shared library with a static CUDA runtime.
"""
from __future__ import annotations

import contextlib
import glob
import hashlib
import os
import shutil
import subprocess
import sys
import tempfile
import warnings
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libttv_b200.so")
STAMP = LIB + ".stamp"            # hash of the sources the library was built from (travels with it, git-ignored like it)
LOCK = os.path.join(PKG, ".build.lock")
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


def sources() -> list[str]:
    """everything the library is compiled from: every file of csrc/ (so that a new header can never be forgotten), the
    public header and this recipe"""
    files = [f for f in glob.glob(os.path.join(CSRC, "*")) if os.path.isfile(f)]
    files += [os.path.join(PKG, "..", "include", "ttv_b200.h"), os.path.abspath(__file__)]
    return sorted(os.path.normpath(f) for f in files)


def source_hash() -> str:
    h = hashlib.sha256()
    for f in sources():
        h.update(os.path.basename(f).encode() + b"\0")
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def stale() -> bool:
    """By CONTENT, not by mtime: a snapshot copied to another box gets fresh mtimes in arbitrary order, and N ranks that all
    decide to rebuild would compile the same files at once."""
    if not os.path.exists(LIB):
        return True
    try:
        with open(STAMP) as f:
            return f.read().strip() != source_hash()
    except OSError:
        return True


@contextlib.contextmanager
def _locked():
    """one builder at a time per tree (torchrun starts N ranks that may all find the library stale)"""
    import fcntl
    fd = os.open(LOCK, os.O_CREAT | os.O_RDWR, 0o644)
    try:
        fcntl.flock(fd, fcntl.LOCK_EX)
        yield
    finally:
        fcntl.flock(fd, fcntl.LOCK_UN)
        os.close(fd)


def _env():
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)       # the image's CC wrapper lacks pieces nvcc's host pass needs
    return env


def _compile(unit, verbose, objdir):
    src, obj, extra = unit
    cmd = [nvcc()] + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", os.path.join(objdir, obj)]
    r = subprocess.run(cmd, env=_env(), capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return r.stderr


def build(force: bool = False, verbose: bool = False, defines=(), out: str = LIB) -> str:
    """defines/out: build an experimental variant beside the product library (e.g. defines=["-DTTVB_MIN_CTAS=4"]).
    Objects go to a private temporary directory and the finished library is renamed into place, under a file lock: ranks
    that start together never see a half-written file, and the second one finds the library fresh."""
    if not force and out == LIB and not stale():
        return LIB
    with _locked():
        if not force and out == LIB and not stale():          # somebody else built it while we waited
            return LIB
        digest = source_hash()
        objdir = tempfile.mkdtemp(prefix="ttv_b200_obj_")
        try:
            todo = [(src, obj, list(extra) + list(defines)) for src, obj, extra in units()]
            with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as pool:
                logs = list(pool.map(lambda u: _compile(u, verbose, objdir), todo))
            if verbose:
                sys.stderr.write("".join(logs))
            tmp_out = os.path.join(objdir, "lib.so")
            cmd = [nvcc()] + ARCH + ["-shared", "--cudart", "static"] + [os.path.join(objdir, u[1]) for u in units()] + ["-o", tmp_out]
            r = subprocess.run(cmd, env=_env(), capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
            staged = out + ".tmp.%d" % os.getpid()
            shutil.copyfile(tmp_out, staged)
            os.chmod(staged, 0o755)
            os.replace(staged, out)                             # atomic on one filesystem
            if out == LIB:
                with open(STAMP + ".tmp", "w") as f:
                    f.write(digest + "\n")
                os.replace(STAMP + ".tmp", STAMP)
        finally:
            shutil.rmtree(objdir, ignore_errors=True)
    return out


def build_or_warn() -> None:
    """what the loader calls: rebuild a stale library; when that is impossible (no nvcc on this box) keep a library that
    travelled with the tree, but say so"""
    try:
        if stale():
            build()
    except Exception as exc:
        if not os.path.exists(LIB):
            raise ImportError(f"ttv_b200: libttv_b200.so is missing and could not be built: {exc}") from exc
        warnings.warn(f"ttv_b200: libttv_b200.so does not match the sources and could not be rebuilt ({str(exc).splitlines()[0]}); "
                      "using the library as it is", RuntimeWarning)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
