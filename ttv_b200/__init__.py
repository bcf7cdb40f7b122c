"""ttv_b200 -- B200-native mode-q tensor-times-vector product behind the tlib::ttv API surface of bassoy/ttv.

    ttv_b200.ttv(q, A, b)               tensor-level interface (numpy = host buffers, torch CUDA tensors = device)
    ttv_b200.ttv_lowlevel(...)          the reference's C-like interface (ttv.h:54-92)
    ttv_b200.ttvpy.ttv / ttvpy.ttvs     drop-in for the reference's Python module
    ttv_b200.sharded                    multi-GPU drivers (one process per GPU, torch.distributed/NCCL)

The arithmetic lives in libttv_b200.so (hand-written sm_100a kernels behind the C-ABI of include/ttv_b200.h).
"""
from .api import (TTVError, Resident, ttv_lowlevel_devices, pinned_empty, ttv, ttv_lowlevel, ttv_multi, ttv_view, ttv_view_scatter, ttv_view_exchange, reduce_slots, plan, plan_view, fill, make_opts, generate_strides,
                  generate_output_shape, generate_output_layout, generate_k_order_layout, is_valid_shape,
                  is_valid_layout, is_valid_strides, launch_count, device_count, DTYPE_CODES)
from . import ttvpy  # noqa: F401

__all__ = ["TTVError", "Resident", "ttv_lowlevel_devices", "pinned_empty", "ttv", "ttv_lowlevel", "ttv_multi", "ttv_view", "ttv_view_scatter", "ttv_view_exchange", "reduce_slots", "plan", "plan_view", "fill", "make_opts",
           "generate_strides", "generate_output_shape", "generate_output_layout", "generate_k_order_layout",
           "is_valid_shape", "is_valid_layout", "is_valid_strides", "launch_count", "device_count", "ttvpy",
           "DTYPE_CODES"]
