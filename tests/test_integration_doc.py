"""INTEGRATION.md section 2 shows the tag and the overload a maintainer of bassoy/ttv would add to route the reference's
own low-level interface into the C-ABI shim.  This test takes that code FROM THE DOCUMENT, compiles it against the
UNMODIFIED reference headers (in place under /root/reference, nothing is copied) and this repo's include/ttv_b200.h, links
libttv_b200.so and calls tlib::ttv::ttv(execution_policy::b200, ...).  Without a GPU the call must surface the shim's
"no CPU fallback" error as the reference's std::runtime_error -- which proves the call went through the shim."""
from __future__ import annotations

import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_INC = "/root/reference/include"

MAIN = r'''
#include <cstdio>
#include <numeric>
#include <vector>
int main()
{
  std::vector<float> a(24), b(3, 1.0f), c(8, 0.0f);
  std::iota(a.begin(), a.end(), 1.0f);
  std::size_t na[] = {4, 3, 2}, wa[] = {1, 4, 12}, pia[] = {1, 2, 3}, nb[] = {3}, nc[] = {4, 2}, wc[] = {1, 4}, pic[] = {1, 2};
  try {
    tlib::ttv::ttv(tlib::ttv::execution_policy::b200, tlib::ttv::slicing_policy::subtensor, tlib::ttv::fusion_policy::all,
                   std::size_t(2), std::size_t(3), a.data(), na, wa, pia, b.data(), nb, c.data(), nc, wc, pic);
    std::printf("RESULT:");
    for (float x : c) std::printf(" %g", x);
    std::printf("\n");
  } catch (std::runtime_error const& e) {
    std::printf("EXC: %s\n", e.what());
  }
  // an invalid argument is reported with the reference's own text, through the same route
  try {
    tlib::ttv::ttv(tlib::ttv::execution_policy::b200, tlib::ttv::slicing_policy::subtensor, tlib::ttv::fusion_policy::all,
                   std::size_t(4), std::size_t(3), a.data(), na, wa, pia, b.data(), nb, c.data(), nc, wc, pic);
  } catch (std::runtime_error const& e) {
    std::printf("EXC2: %s\n", e.what());
  }
  return 0;
}
'''


def test_upstream_snippet_of_integration_md_compiles_and_routes_to_the_shim(tmp_path):
    if not os.path.isdir(REF_INC):
        pytest.skip("the reference tree is not present on this box")
    import ttv_b200
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    section = doc[doc.index("## 2."):doc.index("## 3.")]
    code = re.search(r"```cpp\n(.*?)```", section, flags=re.S).group(1)
    marker = "// include/tlib/detail/tensor_times_vector.h"
    assert marker in code and "// include/tlib/detail/tags.h" in code
    tag_part, overload_part = code[:code.index(marker)], code[code.index(marker):]
    tu = "\n".join([
        "#include <complex>", "#include <type_traits>", "#include <stdexcept>",
        "#include <tlib/detail/tags.h>                  // the reference's, unmodified",
        tag_part,
        "#include <tlib/detail/tensor_times_vector.h>   // the reference's overloads",
        overload_part,
        "#include <tlib/ttv.h>                          // the reference's public interface: its call now finds the new overload",
        MAIN])
    src = tmp_path / "upstream_snippet.cpp"
    src.write_text(tu)
    exe = tmp_path / "upstream_snippet"
    libdir = os.path.join(ROOT, "ttv_b200")
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-DNDEBUG", f"-I{REF_INC}", f"-I{os.path.join(ROOT, 'include')}", str(src), "-o", str(exe),
                        f"-L{libdir}", "-lttv_b200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-4000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    if ttv_b200.device_count() > 0:
        assert "RESULT: 15 18 21 24 51 54 57 60" in r.stdout, r.stdout        # example/interface3.cpp's known answer
    else:
        assert "EXC: Error in ttv_b200: CUDA failure (no CPU fallback exists)." in r.stdout, r.stdout
    assert "EXC2: Error in tlib::tensor_times_vector: contraction mode should be greater zero or less than or equal to p." in r.stdout, r.stdout


def test_ctypes_stub_of_integration_md_calls_the_library():
    """section 4's ctypes stub, taken from the document: argument marshalling must be accepted by the library (without a
    GPU the call returns status 40 = no CPU fallback; the value check of the stub itself runs where a GPU exists)"""
    import sys
    import ttv_b200
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    section = doc[doc.index("## 4."):doc.index("## 5.")]
    code = re.search(r"```python\n(.*?)```", section, flags=re.S).group(1)
    lines = code.rstrip().splitlines()
    assert lines[-1].startswith("assert st == 0")
    if ttv_b200.device_count() == 0:
        lines[-1] = "assert st == 40, st"
    r = subprocess.run([sys.executable, "-c", "\n".join(lines)], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]


def test_pybind_snippets_of_integration_md_build_inside_the_reference_binding(tmp_path):
    """section 3: the two replacements shown for ttvpy/src/wrapped_ttv.cpp are applied to a scratch copy of that file
    (outside the repo), the module is built with pybind11 against the unmodified reference headers and libttv_b200.so, and
    both functions are called: without a GPU they must raise the shim's error as the binding's ValueError"""
    import sys
    import sysconfig
    ref_src = "/root/reference/ttvpy/src/wrapped_ttv.cpp"
    if not os.path.exists(ref_src):
        pytest.skip("the reference tree is not present on this box")
    try:
        import pybind11
    except ImportError:
        pytest.skip("pybind11 is not installed")
    import ttv_b200
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    section = doc[doc.index("## 3."):doc.index("## 4.")]
    blocks = re.findall(r"```cpp\n(.*?)```", section, flags=re.S)
    assert len(blocks) == 2
    ttv_call = blocks[0].split("...\n", 1)[1]                       # after '#include <ttv_b200.h>' and the ellipsis
    ttvs_body = blocks[1]
    text = open(ref_src).read()
    # (1) the three #if branches that pick a CPU policy -> one call into the shim
    i0 = text.index("#ifndef _OPENMP\n    ttv<T>(execution_policy::seq")
    i1 = text.index("#endif", i0) + len("#endif")
    text = text[:i0] + ttv_call + text[i1:]
    # (2) everything in ttvs after the argument checks -> one call of ttv_b200_ttvs
    j0 = text.index("  // B[0]...B[p-2]")
    j1 = text.index("  return c;", j0) + len("  return c;")
    text = text[:j0] + ttvs_body + text[j1:]
    text = "#include <ttv_b200.h>\n" + text
    src = tmp_path / "wrapped_ttv_b200.cpp"
    src.write_text(text)
    out = tmp_path / ("ttvpy" + sysconfig.get_config_var("EXT_SUFFIX"))
    libdir = os.path.join(ROOT, "ttv_b200")
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    cmd = ["g++", "-O1", "-DNDEBUG", "-shared", "-std=c++17", "-fPIC", f"-I{pybind11.get_include()}", f"-I{sysconfig.get_paths()['include']}",
           "-I/root/reference/include", f"-I{os.path.join(ROOT, 'include')}", str(src), "-o", str(out), f"-L{libdir}", "-lttv_b200",
           f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-4000:]
    gpu = ttv_b200.device_count() > 0
    check = f'''
import sys, numpy as np
sys.path.insert(0, {str(tmp_path)!r})
import ttvpy
A = np.arange(24, dtype=np.float64).reshape(3, 2, 4)
gpu = {gpu!r}
for call, want in ((lambda: ttvpy.ttv(1, A, np.arange(3, dtype=np.float64)), np.einsum("ijk,i->jk", A, np.arange(3.0))),
                   (lambda: ttvpy.ttvs(2, A, [np.arange(3.0), np.arange(4.0)], "optimal"), np.einsum("ijk,i,k->j", A, np.arange(3.0), np.arange(4.0)))):
    try:
        got = call()
        assert gpu and np.array_equal(got, want), got
    except ValueError as e:
        assert not gpu and "no CPU fallback" in str(e), e
try:
    ttvpy.ttvs(2, A, [np.arange(3.0)], "optimal")
    raise SystemExit("missing vector was accepted")
except ValueError as e:
    assert "number of input vectors" in str(e)
print("binding ok")
'''
    r = subprocess.run([sys.executable, "-c", check], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "binding ok" in r.stdout, r.stdout + r.stderr[-3000:]


def test_residency_snippet_of_section_1a_compiles_and_runs(tmp_path):
    """tensor::keep_on_device and device_tensor behind the two high-level interfaces -- the
    program in the document, compiled against this repo's include/ and run (on a GPU box it must print the known answer
    of example/interface1.cpp, without one the shim's 'no CPU fallback' error)"""
    import ttv_b200
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    section = doc[doc.index("### 1a."):doc.index("## 2.")]
    code = re.search(r"```cpp\n(.*?)```", section, flags=re.S).group(1)
    assert "keep_on_device" in code and "device_tensor" in code
    src = tmp_path / "residency.cpp"
    src.write_text(code)
    exe = tmp_path / "residency"
    libdir = os.path.join(ROOT, "ttv_b200")
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", f"-I{os.path.join(ROOT, 'include')}", str(src), "-o", str(exe),
                        f"-L{libdir}", "-lttv_b200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-4000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    if ttv_b200.device_count() > 0:
        assert "C1: 15 18 21 24 51 54 57 60" in r.stdout and "C2[0] 15 C3[0] 6 C4[0] 6" in r.stdout, r.stdout
    else:
        assert "EXC: Error in ttv_b200: CUDA failure (no CPU fallback exists)." in r.stdout, r.stdout


@pytest.mark.gpu
def test_residency_snippet_of_section_1a_on_the_gpu(tmp_path):
    """the same program on a B200: it must print example/interface1.cpp's known answer through all four routes"""
    test_residency_snippet_of_section_1a_compiles_and_runs(tmp_path)
