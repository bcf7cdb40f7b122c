// scatter_kernel.cuh -- the n_q-split product of the multi-GPU path FUSED with its exchange step.
//
// When mode q is the slowest mode of the layout, G GPUs each hold a contiguous range of the contraction rows and compute
// a full-size PARTIAL of C; the partials have to be summed ("one ncclReduce").  Here the exchange rides on the
// kernel's own stores: the flat index space of C is cut into G blocks of `blk` elements, block j belongs to GPU j, and
// every GPU writes the piece of its partial that falls into block j straight into GPU j's memory -- slot `rank` of
// j's workspace [G][blk], mapped into this process over NVLink / NVSwitch (peer pointers from symmetric memory).  The
// 16-byte stores leave while the kernel is still streaming A, so the transfer (|C|·(G-1)/G bytes per GPU) hides behind
// the HBM-bound main loop.  After a cross-GPU barrier each GPU sums its G slots in rank order (ttv_reduce_kernel:
// deterministic) and holds its block of C: a reduce-scatter, the same distributed form the free-split products leave C in.
//
// The arithmetic is the column GEMV of kernels.cuh (reference gemv_col, detail/matrix_times_vector.h:108-127) with a
// whole CTA along inner and eight k-steps in flight.
#pragma once

#include "kernels.cuh"

namespace ttvb {

constexpr int kMaxPeers = 16;

struct ScatterParams {
  const void* a;
  const void* b;
  void*       peer[kMaxPeers];   // peer[j]: GPU j's workspace [world][blk], addressable from this GPU
  uint64_t outer, nq, inner;
  uint64_t blk;                  // elements of C's flat index space per GPU (multiple of the vector width)
  uint64_t itiles, tiles;
  uint32_t world, rank;
  uint32_t kb;                   // elements of b per shared-memory chunk
  uint32_t stream;
};

template<class T, int V>
__global__ void __launch_bounds__(256, 3)
ttv_col_scatter_kernel(const ScatterParams P)
{
  constexpr int KU = 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sb = reinterpret_cast<T*>(smem_raw);            // [kb]

  const T* __restrict__ A = static_cast<const T*>(P.a);
  const T* __restrict__ B = static_cast<const T*>(P.b);
  const uint32_t tid = threadIdx.x;
  const bool stream = P.stream != 0;
  const bool b_resident = P.nq <= P.kb;

  if (b_resident) {
    for (uint32_t j = tid; j < (uint32_t)P.nq; j += blockDim.x) sb[j] = B[j];
    __syncthreads();
  }

  for (uint64_t tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
    const uint64_t it = tile % P.itiles;
    const uint64_t o  = tile / P.itiles;
    const uint64_t i0 = (it * blockDim.x + tid) * V;
    const int nvalid = i0 < P.inner ? 1 : 0;

    T acc[1][V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[0][j] = Num<T>::zero();

    for (uint64_t k0 = 0; k0 < P.nq; k0 += P.kb) {
      const uint32_t kn = (uint32_t)min((uint64_t)P.kb, P.nq - k0);
      if (!b_resident) {
        __syncthreads();
        for (uint32_t j = tid; j < kn; j += blockDim.x) sb[j] = B[k0 + j];
        __syncthreads();
      }
      if (nvalid) {
        const T* ap = A + (o * P.nq + k0) * P.inner + i0;
        uint32_t k = 0;
        for (; k + (KU - 1) < kn; k += KU, ap += KU * P.inner)
          col_batch<T, V, 1, KU, false>(acc, ap, 0, P.inner, sb, k, 1u, kn, 1, stream);
        for (; k < kn; k += KU, ap += KU * P.inner)
          col_batch<T, V, 1, KU, true>(acc, ap, 0, P.inner, sb, k, 1u, kn, 1, stream);
      }
    }

    if (nvalid) {
      // the owner of this vector of C and the slot this GPU writes there
      const uint64_t f = o * P.inner + i0;
      const uint64_t j = f / P.blk;
      T* dst = static_cast<T*>(P.peer[j]) + (uint64_t)P.rank * P.blk + (f - j * P.blk);
      Vec<T, V> val;
#pragma unroll
      for (int e = 0; e < V; ++e) val.e[e] = acc[0][e];
      *reinterpret_cast<Vec<T, V>*>(dst) = val;
    }
  }
}

} // namespace ttvb
