"""The device-residency and multi-device entries of the C-ABI (include/ttv_b200.h, version 110):

  ttv_b200_run_resident   a HOST tensor keeps its copy in HBM between products (what tlib::ttv::tensor::keep_on_device uses)
  ttv_b200_run_devices    one host tensor cut along its slowest mode over several GPUs of the process
  ttv_b200_copy / ttv_b200_host_alloc / ttv_b200_device_alloc   memory for header-style hosts (tlib::ttv::device_tensor)

Every result is compared with the oracle bit for bit (integer-valued data) or within the stated tolerance."""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

import ttv_b200
from conftest import assert_close, random_case, real_case

pytestmark = pytest.mark.gpu


def lowlevel_args(q, a, na, pia, b):
    nc = ttv_b200.generate_output_shape(na, q); pic = ttv_b200.generate_output_layout(pia, q)
    c = np.full(a.size // na[q - 1], 77, a.dtype)
    return (q, len(na), a, list(na), ttv_b200.generate_strides(na, pia), list(pia), b, [len(b)], c, nc, ttv_b200.generate_strides(nc, pic), pic), c


CASES = [((37, 21, 40), (1, 2, 3)), ((16, 33, 9, 12), (2, 1, 4, 3)), ((300, 17), (1, 2)), ((17, 300), (2, 1)), ((5, 6, 7, 19), (4, 3, 2, 1))]


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.complex64])
def test_resident_tensor_uploads_once_and_follows_invalidation(oracle, dtype, monkeypatch):
    rng = np.random.default_rng(5)
    for chunk_mb in ("128", "1"):                                  # plain upload / chunks streamed under the first product
        monkeypatch.setenv("TTV_B200_H2D_CHUNK_MB", chunk_mb)
        for na, pia in CASES + [((64, 50, 203), (1, 2, 3)), ((211, 3000), (2, 1))]:
            a, _ = random_case(rng, na, 1, dtype)
            res = ttv_b200.Resident()
            assert not res.valid
            for q in range(1, len(na) + 1):                        # every mode of ONE tensor: the benchmark protocol
                _, b = random_case(rng, na, q, dtype)
                args, c = lowlevel_args(q, a, na, pia, b)
                res.ttv_lowlevel(*args)
                assert res.valid
                assert np.array_equal(c, oracle.ttv(q, a, na, pia, b)), (na, pia, q, chunk_mb)
            # the host data changes: without invalidate() the OLD copy answers, with it the new data does
            a_old = a.copy()
            a[:] = a[::-1].copy()
            _, b = random_case(rng, na, 1, dtype)
            args, c = lowlevel_args(1, a, na, pia, b)
            res.ttv_lowlevel(*args)
            assert np.array_equal(c, oracle.ttv(1, a_old, na, pia, b))
            res.invalidate()
            assert not res.valid
            res.ttv_lowlevel(*args)
            assert np.array_equal(c, oracle.ttv(1, a, na, pia, b))
            # another array through the same twin: recognised by its address / size, uploaded again
            a3, b3 = random_case(rng, na[:2] if len(na) > 2 else na, 1, dtype)
            na3 = na[:2] if len(na) > 2 else na
            args, c = lowlevel_args(1, a3, na3, (1, 2), b3)
            res.ttv_lowlevel(*args)
            assert np.array_equal(c, oracle.ttv(1, a3, na3, (1, 2), b3))
            res.close()


def test_resident_keeps_error_behaviour_and_accumulate(oracle):
    rng = np.random.default_rng(6)
    na, pia = (12, 9, 14), (2, 3, 1)
    a, b = random_case(rng, na, 2, np.float64)
    res = ttv_b200.Resident()
    args, c = lowlevel_args(2, a, na, pia, b)
    with pytest.raises(ttv_b200.TTVError) as err:
        res.ttv_lowlevel(*((5,) + args[1:]))
    assert err.value.status == 2
    assert not res.valid
    c[:] = 3
    res.ttv_lowlevel(*args, flags=ttv_b200.api.FLAG_ACCUMULATE)
    assert np.array_equal(c, oracle.ttv(2, a, na, pia, b) + 3)
    c[:] = 5
    res.ttv_lowlevel(*args, flags=ttv_b200.api.FLAG_ACCUMULATE)    # second product: from HBM
    assert np.array_equal(c, oracle.ttv(2, a, na, pia, b) + 5)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.complex64])
def test_one_host_tensor_over_several_devices(oracle, dtype, monkeypatch):
    """free split and n_q split; a single-GPU box lists its device several times (two host threads, serialised inside the
    library), a multi-GPU box uses every device.  Uneven splits (37 over 3), more devices than slabs (19 over 4 -> 5)."""
    import torch
    n_dev = torch.cuda.device_count()
    monkeypatch.setenv("TTV_B200_MULTI_MIN_MB", "0")                # these tensors are tiny: split them anyway
    monkeypatch.setenv("TTV_B200_MULTI_PAGEABLE_NQ", "1")           # ... also the n_q split of these PAGEABLE arrays (kept on one device by default)
    rng = np.random.default_rng(8)
    for devices in ([0, 0], list(range(n_dev)) if n_dev > 1 else [0, 0, 0], [0] * 5, [0]):
        for chunk_mb in ("128", "1"):
            monkeypatch.setenv("TTV_B200_H2D_CHUNK_MB", chunk_mb)
            for na, pia in CASES:
                for q in range(1, len(na) + 1):
                    a, b = random_case(rng, na, q, dtype)
                    args, c = lowlevel_args(q, a, na, pia, b)
                    ttv_b200.ttv_lowlevel_devices(devices, *args)
                    assert np.array_equal(c, oracle.ttv(q, a, na, pia, b)), (devices, na, pia, q, chunk_mb)
    # accumulate through the n_q split (C goes up, the partials are added to it)
    na, pia = (300, 17), (2, 1)
    a, b = random_case(rng, na, 1, dtype)
    args, c = lowlevel_args(1, a, na, pia, b)
    c[:] = 2
    ttv_b200.ttv_lowlevel_devices([0, 0, 0], *args, flags=ttv_b200.api.FLAG_ACCUMULATE)
    assert np.array_equal(c, oracle.ttv(1, a, na, pia, b) + 2)
    with pytest.raises(ttv_b200.TTVError):
        ttv_b200.ttv_lowlevel_devices([0, 99], *args)


def test_devices_on_real_valued_data_and_pinned_memory(oracle, monkeypatch):
    monkeypatch.setenv("TTV_B200_MULTI_MIN_MB", "0")
    rng = np.random.default_rng(9)
    na, pia = (96, 80, 130), (1, 2, 3)
    for q in (1, 2, 3):
        a0, b = real_case(rng, na, q, np.float32)
        a = ttv_b200.pinned_empty(a0.size, np.float32)
        a[:] = a0
        args, c = lowlevel_args(q, a, na, pia, b)
        ttv_b200.ttv_lowlevel_devices([0, 0, 0], *args)
        ref, mag = oracle.naive(q, a0, na, pia, b, want_abs=True)
        assert_close(c, ref, mag, na[q - 1], np.float32, what=f"q={q}")


def test_copy_and_memory_helpers_round_trip():
    lib = ttv_b200._lib.load()
    n = (24 << 20) + 12345                                          # > 8 MiB of pageable memory: the pipelined bounce path
    src = np.random.default_rng(1).integers(0, 255, n).astype(np.uint8)
    dev = C.c_void_p()
    assert lib.ttv_b200_device_alloc(C.byref(dev), n, -1, 1) == 0
    back = np.full(n, 9, np.uint8)
    assert lib.ttv_b200_copy(back.ctypes.data_as(C.c_void_p), dev, n, None) == 0 and not back.any()          # zero-initialised
    assert lib.ttv_b200_copy(dev, src.ctypes.data_as(C.c_void_p), n, None) == 0
    assert lib.ttv_b200_copy(back.ctypes.data_as(C.c_void_p), dev, n, None) == 0
    assert np.array_equal(back, src)
    pinned = ttv_b200.pinned_empty(n, np.uint8)
    assert lib.ttv_b200_copy(pinned.ctypes.data_as(C.c_void_p), dev, n, None) == 0
    assert np.array_equal(pinned, src)
    dev2 = C.c_void_p()
    assert lib.ttv_b200_device_alloc(C.byref(dev2), n, 0, 0) == 0
    assert lib.ttv_b200_copy(dev2, dev, n, None) == 0
    back[:] = 0
    assert lib.ttv_b200_copy(back.ctypes.data_as(C.c_void_p), dev2, n, None) == 0
    assert np.array_equal(back, src)
    assert lib.ttv_b200_device_free(dev) == 0 and lib.ttv_b200_device_free(dev2) == 0


def test_resident_and_devices_with_padded_strides(oracle, monkeypatch):
    """a tensor with padded leading dimensions (TTV_B200_FLAG_HONOR_STRIDES) through the resident twin and through a device
    list: the whole span is what lives on the device / what a single device gets (strided views are not cut)"""
    monkeypatch.setenv("TTV_B200_MULTI_MIN_MB", "0")
    rng = np.random.default_rng(12)
    na, pia = (10, 7, 9), (1, 2, 3)
    wa = [1, 12, 12 * 8]                                             # rows padded 10 -> 12, slabs 7 -> 8 rows
    span = 1 + sum((n - 1) * w for n, w in zip(na, wa))
    buf = np.full(span, 77, np.float64)
    x = rng.integers(-8, 9, na).astype(np.float64)
    idx = np.add.outer(np.add.outer(np.arange(10) * wa[0], np.arange(7) * wa[1]), np.arange(9) * wa[2])
    buf[idx] = x
    packed = np.ascontiguousarray(x.transpose(2, 1, 0)).reshape(-1)   # the same tensor, packed first-order
    res = ttv_b200.Resident()
    for q in (1, 2, 3):
        b = rng.integers(-8, 9, na[q - 1]).astype(np.float64)
        want = oracle.ttv(q, packed, na, pia, b)
        nc = ttv_b200.generate_output_shape(na, q); pic = ttv_b200.generate_output_layout(pia, q)
        for call in (lambda *a, **k: res.ttv_lowlevel(*a, **k), lambda *a, **k: ttv_b200.ttv_lowlevel_devices([0, 0], *a, **k)):
            c = np.full(want.size, 5, np.float64)
            call(q, 3, buf, list(na), wa, list(pia), b, [len(b)], c, nc, ttv_b200.generate_strides(nc, pic), pic, flags=8)
            assert np.array_equal(c, want), (q, call)
    assert res.valid


def test_copy_between_host_buffers_uses_the_copy_threads():
    lib = ttv_b200._lib.load()
    n = (40 << 20) + 7
    src = np.random.default_rng(2).integers(0, 255, n).astype(np.uint8)
    dst = np.zeros(n, np.uint8)
    assert lib.ttv_b200_copy(dst.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p), n, None) == 0
    assert np.array_equal(dst, src)
