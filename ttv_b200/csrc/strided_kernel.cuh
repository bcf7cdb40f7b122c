// strided_kernel.cuh -- STRIDED: the general-stride form of the product, for tensors whose strides are NOT the packed
// strides of their shape and layout (padded leading dimensions, views into larger arrays).
//
// The reference honours such strides only in its slice variants, where the loop nest advances a and c with wa / wc per
// free mode (detail/tensor_times_vector.h:189-216: a + i*wa[pia[r-1]-1], c + i*wc[pic[q-1]-1]) and the leaf GEMV runs
// with lda = wa[q-1] (:214, matrix_times_vector.h:108-127).  Here the free modes are listed in the order of C's layout
// (fastest first), neighbours that are packed against each other in BOTH tensors are folded into one, and one thread
// owns one output: it decodes its mixed-radix index into the element offsets of A and C and walks n_q with stride
// wa[q-1], eight independent loads in flight.  Consecutive threads run along C's fastest mode, so the loads coalesce
// whenever that mode has stride 1 in A.  This is a correctness-first path; the packed kernels are the fast ones.
#pragma once

#include "numeric.cuh"

namespace ttvb {

constexpr int kMaxFree = 8;

struct StridedParams {
  const void* a;
  const void* b;
  void*       c;
  uint64_t nq, wq;          // contraction extent and its stride in A
  uint64_t total;           // outputs
  uint64_t n[kMaxFree], wa[kMaxFree], wc[kMaxFree];
  uint32_t nfree;
  uint32_t accumulate;
};

template<class T>
__global__ void __launch_bounds__(256)
ttv_strided_kernel(const StridedParams P)
{
  const T* __restrict__ A = static_cast<const T*>(P.a);
  const T* __restrict__ B = static_cast<const T*>(P.b);
  T* __restrict__       C = static_cast<T*>(P.c);
  constexpr int KU = 8;

  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < P.total; j += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t rem = j, offa = 0, offc = 0;
    for (uint32_t d = 0; d < P.nfree; ++d) {
      const uint64_t i = rem % P.n[d];
      rem /= P.n[d];
      offa += i * P.wa[d];
      offc += i * P.wc[d];
    }
    const T* ap = A + offa;
    T acc = Num<T>::zero();
    uint64_t k = 0;
    for (; k + KU <= P.nq; k += KU) {
      T v[KU];
#pragma unroll
      for (int s = 0; s < KU; ++s) v[s] = ap[(k + s) * P.wq];
#pragma unroll
      for (int s = 0; s < KU; ++s) acc = Num<T>::madd(v[s], B[k + s], acc);
    }
    for (; k < P.nq; ++k) acc = Num<T>::madd(ap[k * P.wq], B[k], acc);
    T* out = C + offc;
    *out = P.accumulate ? Num<T>::add(*out, acc) : acc;
  }
}

} // namespace ttvb
