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
// The whole exchange in ONE kernel (ttv_col_exchange_kernel): product + scatter, cross-GPU flag barrier, sum of the slots into
// this GPU's block of C.  flags[j]: GPU j's flag array; counter / error: 8 + 4 bytes of local device memory (zero at first use).
// max_ctas caps the CTAs that stay for the barrier and the slot sum (0 = one per SM).
cudaError_t launch_exchange(int dtype, const View& v, const void* a, const void* b, void* const* peers, void* const* flags,
                            uint32_t world, uint32_t rank, uint64_t blk, int vec, void* c_block, uint64_t n_block, uint32_t token,
                            void* counter, void* error, bool accumulate, uint64_t timeout_ns, uint32_t max_ctas, int sm_count,
                            cudaStream_t stream);
cudaError_t launch_reduce_slots(int dtype, const void* ws, void* c, uint64_t n, uint64_t stride, uint32_t slots, bool accumulate,
                                int sm_count, cudaStream_t stream);

// x[i] = synth(seed, first + i) for i < count, on the device (same generator as oracle/ttv_oracle.c).
cudaError_t launch_fill(int dtype, void* x, uint64_t first, uint64_t count, uint64_t seed, int sm_count, cudaStream_t stream);

uint64_t launch_count();

} // namespace ttvb
