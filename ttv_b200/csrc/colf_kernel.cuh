// colf_kernel.cuh -- COLF: column GEMV whose rows are NARROWER than / not a multiple of a 16-byte vector, read as a flat
// stream of whole vectors (sm_100a).
//
// A[outer][n_q][inner] with inner = 2, 3, 5, 6, 7, 9 ... elements under a long contraction: the asymmetric family of the
// reference (tiny leading extents, one huge mode; its slicing::small leaf is an n1 x n_q GEMV, tensor_times_vector.h:214,
// gemv_col matrix_times_vector.h:108-127).  No 16-byte vector tiles such a row, so the column kernel loads 4 or 8 bytes
// per lane (4.2-4.3 TB/s for rows of 3 / 5 / 7 floats).  But R = V / gcd(inner, V) consecutive rows -- a SUPER-ROW -- are
// L = inner / gcd(inner, V) whole vectors, and a slab is a contiguous run of super-rows.  So the CTA streams the slab flat:
// thread (ty, j) loads vector j of super-rows ty, ty + TY, ... (consecutive threads read consecutive vectors: a warp reads
// 512 contiguous bytes per instruction, KU loads in flight per thread).  Element e of that vector is the element
// (row, column) = ((V j + e) div inner, (V j + e) mod inner) of its super-row -- the same for every super-row the thread
// visits -- so the thread keeps V accumulators, one per element, and multiplies element e by b[R sr + row_e]: the R
// elements of b of a super-row are one aligned 8- / 16-byte load that the L threads of the super-row share (L1 hit), and
// row_e selects among them with compile-time-unrolled selects.  Afterwards the partial sums of a CTA, written to shared
// memory in thread order, ARE a matrix [TY R][inner] whose columns are summed by a flat tree (fixed order: deterministic).
// A contraction cut across CTAs (ksplit) goes to the workspace [ksplit][outer * inner] like in the column kernel.
#pragma once

#include "kernels.cuh"

namespace ttvb {

struct ColfParams {
  const void* a;
  const void* b;
  void*       c;            // C, or the workspace [ksplit][outer * inner] when ksplit > 1
  uint64_t outer, nq, inner;
  uint64_t srchunk;         // super-rows per partition
  uint32_t ksplit;
  uint32_t L, TY;           // vectors per super-row, super-rows per step of the CTA (TY * L <= 256 threads work)
  uint32_t accumulate;      // only honoured when ksplit == 1
};

template<class T, int R>
__device__ __forceinline__ T colf_select(const Vec<T, R>& q, uint32_t r)
{
  if constexpr (R == 2) return r ? q.e[1] : q.e[0];
  else                  return r < 2 ? (r == 0 ? q.e[0] : q.e[1]) : (r == 2 ? q.e[2] : q.e[3]);
}

template<class T, int R, int KU, bool PRED>
__device__ __forceinline__ void colf_batch(T (&acc)[16 / sizeof(T)], const T* ap, const T* bp, uint64_t astep, uint64_t bstep,
                                           uint64_t sr, uint64_t step, uint64_t n, const uint32_t (&row)[16 / sizeof(T)])
{
  constexpr int V = 16 / (int)sizeof(T);
  Vec<T, V> x[KU];
  Vec<T, R> q[KU];
#pragma unroll
  for (int s = 0; s < KU; ++s) {
    if constexpr (PRED) {
      const bool ok = sr + s * step < n;
      x[s] = ok ? load_a<T, V>(ap + s * astep, true) : zero_vec<T, V>();
      q[s] = ok ? *reinterpret_cast<const Vec<T, R>*>(bp + s * bstep) : zero_vec<T, R>();
    } else {
      x[s] = load_a<T, V>(ap + s * astep, true);
      q[s] = *reinterpret_cast<const Vec<T, R>*>(bp + s * bstep);
    }
  }
#pragma unroll
  for (int s = 0; s < KU; ++s)
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = Num<T>::madd(x[s].e[e], colf_select<T, R>(q[s], row[e]), acc[e]);
}

template<class T, int R, int KU>
__global__ void __launch_bounds__(256, 3)
ttv_colf_kernel(const ColfParams P)
{
  pdl_prologue();
  constexpr int V = 16 / (int)sizeof(T);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* red = reinterpret_cast<T*>(smem_raw);                          // [TY * L * V] = [TY * R][inner]

  const T* __restrict__ A = static_cast<const T*>(P.a);
  const T* __restrict__ B = static_cast<const T*>(P.b);
  T* __restrict__       C = static_cast<T*>(P.c);

  const uint32_t tid = threadIdx.x;
  const uint32_t inner = (uint32_t)P.inner;
  const uint32_t j  = tid % P.L;
  const uint32_t ty = tid / P.L;
  const bool live = ty < P.TY;
  uint32_t row[V];
#pragma unroll
  for (int e = 0; e < V; ++e) row[e] = (V * j + e) / inner;

  const uint64_t nsr   = P.nq / R;                                  // whole super-rows of a slab
  const uint32_t work  = P.TY * P.L;                                // threads that take part = vectors per step
  const uint32_t cells = work * V;                                  // partial sums of the CTA
  const uint64_t astep = (uint64_t)work * V, bstep = (uint64_t)P.TY * R;
  const uint64_t items = P.outer * P.ksplit;

  for (uint64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const uint64_t o  = item / P.ksplit;
    const uint32_t ks = (uint32_t)(item % P.ksplit);
    const uint64_t srbeg = min((uint64_t)ks * P.srchunk, nsr), srend = min(srbeg + P.srchunk, nsr);
    const uint64_t n = srend - srbeg;

    T acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = Num<T>::zero();

    if (live) {
      const T* ap = A + (o * P.nq + srbeg * R) * inner + (uint64_t)tid * V;
      const T* bp = B + (srbeg + ty) * R;
      uint64_t sr = ty;
      for (; sr + (uint64_t)(KU - 1) * P.TY < n; sr += (uint64_t)KU * P.TY, ap += KU * astep, bp += KU * bstep)
        colf_batch<T, R, KU, false>(acc, ap, bp, astep, bstep, sr, P.TY, n, row);
      if (sr < n) colf_batch<T, R, KU, true>(acc, ap, bp, astep, bstep, sr, P.TY, n, row);
    }

    __syncthreads();                                                // red may still be read by the item before
    if (live) {
#pragma unroll
      for (int e = 0; e < V; ++e) red[tid * V + e] = acc[e];
    }
    __syncthreads();
    // column sums of the matrix [TY * R][inner] that red now is: flat halving tree over whole rows
    const uint32_t rows = P.TY * R;
    uint32_t span = 1;
    while (span < rows) span <<= 1;
    for (uint32_t h = span >> 1; h > 0; h >>= 1) {
      const uint32_t lim = min(h * inner, cells > h * inner ? cells - h * inner : 0u);
      for (uint32_t idx = tid; idx < lim; idx += blockDim.x) red[idx] = Num<T>::add(red[idx], red[idx + h * inner]);
      __syncthreads();
      // rows [h, 2h) are folded into [0, h); what lay beyond 2h does not exist (span is the next power of two)
    }
    if (tid < inner) {
      T val = red[tid];
      if (ks + 1 == P.ksplit) {                                     // rows past the last whole super-row (only when outer == 1)
        for (uint64_t r = nsr * R; r < P.nq; ++r) val = Num<T>::madd(A[(o * P.nq + r) * inner + tid], B[r], val);
      }
      T* out = C + ((P.ksplit > 1 ? (uint64_t)ks * P.outer : 0) + o) * inner + tid;
      *out = (P.accumulate && P.ksplit == 1) ? Num<T>::add(*out, val) : val;
    }
  }
}

} // namespace ttvb
