// dotp_kernel.cuh -- DOTP: fibers of TWO elements (inner = 1, n_q = 2), 4- and 8-byte element types (sm_100a).
//
// The asymmetric family of the reference contracts tiny leading modes of huge tensors ([2, 2, 4, 2, 2^15, ...], q = 1:
// 1.6 * 10^9 fibers of two elements; its slicing::small leaf is a 2-element dot, detail/matrix_times_vector.h:131-148 under
// the loop nest detail/tensor_times_vector.h:189-324).  A fiber is narrower than (4-byte types) or exactly (8-byte types) one
// 16-byte vector, a third of the traffic is the WRITE of C, and there is nothing to reduce across threads: a 16-byte vector
// of A holds whole fibers, so a lane loads consecutive vectors (a warp reads 512 contiguous bytes per instruction, KU of them
// in flight per lane), multiplies by the two elements of b it keeps in registers and stores the 8 bytes of C that belong to
// its vector (a warp writes 256 contiguous bytes per instruction).  No shared memory, no synchronisation, ONE tile of
// 256 x KU vectors per CTA: CTAs striding over the tiles measured 6 432-6 478 GB/s on [1610612736, 2, 1], a CTA per tile
// 7 002-7 060 (STREAM through shared memory: 6 311-6 539; 8-byte types 6 957-6 970 against DOTF's 6 689-6 738).  Swapping
// halves between neighbouring lanes for 16-byte stores measured slower (6 692) and is gone.
#pragma once

#include "numeric.cuh"

namespace ttvb {

struct DotpParams {
  const void* a;
  const void* b;
  void*       c;
  uint64_t outer;           // fibers
  uint64_t nvec;            // whole 16-byte vectors of A: floor(outer * 2 * sizeof(T) / 16)
  uint64_t tiles;           // ceil(nvec / (256 * KU))
  uint32_t accumulate;
};

template<class T, int KU>
__global__ void __launch_bounds__(256)
ttv_dotp_kernel(const DotpParams P)
{
  pdl_prologue();
  constexpr int V  = 16 / (int)sizeof(T);       // elements per vector of A
  constexpr int OV = V / 2;                     // outputs per vector: 8 bytes
  using VA = Vec<T, V>;
  using VC = Vec<T, OV>;
  const T* __restrict__ A = static_cast<const T*>(P.a);
  const T* __restrict__ B = static_cast<const T*>(P.b);
  T* __restrict__       C = static_cast<T*>(P.c);
  const T b0 = B[0], b1 = B[1];

  auto product = [&](const VA& x, uint64_t vec) {
    VC y;
#pragma unroll
    for (int i = 0; i < OV; ++i) y.e[i] = Num<T>::madd(x.e[2 * i + 1], b1, Num<T>::madd(x.e[2 * i], b0, Num<T>::zero()));
    VC* out = reinterpret_cast<VC*>(C + vec * OV);
    if (P.accumulate) {
      const VC old = *out;
#pragma unroll
      for (int i = 0; i < OV; ++i) y.e[i] = Num<T>::add(old.e[i], y.e[i]);
    }
    *out = y;
  };

  for (uint64_t tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
    const uint64_t v0 = tile * (uint64_t)(256 * KU) + threadIdx.x;
    VA x[KU];
    if (v0 - threadIdx.x + (uint64_t)(256 * KU) <= P.nvec) {
#pragma unroll
      for (int j = 0; j < KU; ++j) x[j] = load_stream<T, V>(A + (v0 + (uint64_t)j * 256) * V);
#pragma unroll
      for (int j = 0; j < KU; ++j) product(x[j], v0 + (uint64_t)j * 256);
    } else {
#pragma unroll
      for (int j = 0; j < KU; ++j)
        if (v0 + (uint64_t)j * 256 < P.nvec) x[j] = load_stream<T, V>(A + (v0 + (uint64_t)j * 256) * V);
#pragma unroll
      for (int j = 0; j < KU; ++j)
        if (v0 + (uint64_t)j * 256 < P.nvec) product(x[j], v0 + (uint64_t)j * 256);
    }
  }

  // an odd number of 4-byte fibers leaves half a vector: one fiber, plain loads
  if (blockIdx.x == 0 && threadIdx.x == 0 && P.nvec * OV < P.outer) {
    const uint64_t f = P.outer - 1;
    T y = Num<T>::madd(A[2 * f + 1], b1, Num<T>::madd(A[2 * f], b0, Num<T>::zero()));
    C[f] = P.accumulate ? Num<T>::add(C[f], y) : y;
  }
}

} // namespace ttvb
