// launch.h -- kernel launch entry points (implemented in launch.cu).
#pragma once

#include <cuda_runtime.h>
#include "plan.h"

namespace ttvb {

// Runs one TTV on the canonical view with DEVICE pointers.  `workspace` must hold l.workspace_bytes when l.ksplit > 1.
cudaError_t launch_view(int dtype, const View& v, const Launch& l, const void* a, const void* b, void* c,
                        void* workspace, bool accumulate, int sm_count, cudaStream_t stream);

// General strides (v.strided): one thread per output, DEVICE pointers (strided_kernel.cuh).
cudaError_t launch_strided(int dtype, const View& v, const void* a, const void* b, void* c, bool accumulate, int sm_count,
                           cudaStream_t stream);

// Fused n_q-split product + exchange over peer memory (scatter_kernel.cuh) and the sum of the received slots.
cudaError_t launch_scatter(int dtype, const View& v, const void* a, const void* b, void* const* peers, uint32_t world, uint32_t rank,
                           uint64_t blk, int vec, int sm_count, cudaStream_t stream);
cudaError_t launch_reduce_slots(int dtype, const void* ws, void* c, uint64_t n, uint64_t stride, uint32_t slots, bool accumulate,
                                int sm_count, cudaStream_t stream);

// x[i] = synth(seed, first + i) for i < count, on the device (same generator as oracle/ttv_oracle.c).
cudaError_t launch_fill(int dtype, void* x, uint64_t first, uint64_t count, uint64_t seed, int sm_count, cudaStream_t stream);

uint64_t launch_count();

} // namespace ttvb
