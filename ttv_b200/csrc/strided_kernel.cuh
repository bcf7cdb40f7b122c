// strided_kernel.cuh -- STRIDED: the general-stride form of the product, for tensors whose strides are NOT the packed
// strides of their shape and layout (padded leading dimensions, views into larger arrays).
//
// The reference honours such strides only in its slice variants, where the loop nest advances a and c with wa / wc per
// free mode (detail/tensor_times_vector.h:189-216: a + i*wa[pia[r-1]-1], c + i*wc[pic[q-1]-1]) and the leaf GEMV runs
// with lda = wa[q-1] (:214, matrix_times_vector.h:108-127).  Here the free modes are listed in the order of C's layout
// (fastest first), neighbours that are packed against each other in BOTH tensors are folded into one, and one thread
// owns one output: it decodes its mixed-radix index into the element offsets of A and C and walks n_q with stride
// wa[q-1], eight independent loads in flight.  Consecutive threads run along C's fastest mode, so the loads coalesce
// whenever that mode has stride 1 in A (measured on slices of a 256^4 fp32 tensor, q = 2..4: 6.1-6.4 TB/s).  When q
// itself is the contiguous mode the lanes would sit on 32 different fibers: ttv_strided_dot_kernel below takes those.
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
  uint64_t n0p = 0;         // vector form: lanes per row of the fastest free mode, n[0]/V rounded up to whole warps
};

// output index j -> element offsets of A and C: mixed-radix decode over the free modes, fastest first.  32-bit
// division when the number of outputs allows (a 64-bit division costs ~100 instructions, and short fibers pay one
// decode per handful of loads).
__device__ __forceinline__ void strided_decode(const StridedParams& P, uint64_t j, uint64_t& offa, uint64_t& offc)
{
  offa = 0; offc = 0;
  if (P.total <= 0xffffffffull) {
    uint32_t rem = (uint32_t)j;
    for (uint32_t d = 0; d < P.nfree; ++d) {
      const uint32_t nd = (uint32_t)P.n[d];
      const uint32_t quo = rem / nd, i = rem - quo * nd;
      rem = quo;
      offa += (uint64_t)i * P.wa[d];
      offc += (uint64_t)i * P.wc[d];
    }
  } else {
    uint64_t rem = j;
    for (uint32_t d = 0; d < P.nfree; ++d) {
      const uint64_t i = rem % P.n[d];
      rem /= P.n[d];
      offa += i * P.wa[d];
      offc += i * P.wc[d];
    }
  }
}

template<class T>
__global__ void __launch_bounds__(256)
ttv_strided_kernel(const StridedParams P)
{
  pdl_prologue();
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

// ------------------------------------------------------------------------------------------------------------------
// STRIDED, vector form: the fastest free mode is contiguous in A AND in C and every other stride, the extent of that
// mode and the addresses are multiples of V elements (a leading dimension padded to whole 16-byte lines, a slice that
// cuts a slower mode).  A thread then owns V CONSECUTIVE outputs: one 16-byte load per k-step instead of V scalar ones,
// eight of them (128 bytes) in flight, one 16-byte store; the index decode is shared by the V outputs.
// Rows are handed out in whole warps (n0p = n[0]/V rounded up to a multiple of 32 lanes, the surplus lanes idle): a warp
// whose 32 vectors straddle two rows of a padded tensor measured 7 % MORE DRAM reads than the tensor holds (ncu on rows
// of 248 of 256 floats: 18.44 GB against 17.25 GB for the thread-per-output form) and lost to it.
// ------------------------------------------------------------------------------------------------------------------
template<class T, int V>
__global__ void __launch_bounds__(256)
ttv_strided_vec_kernel(const StridedParams P)
{
  pdl_prologue();
  const T* __restrict__ A = static_cast<const T*>(P.a);
  const T* __restrict__ B = static_cast<const T*>(P.b);
  T* __restrict__       C = static_cast<T*>(P.c);
  constexpr int KU = 8;
  const uint64_t n0v = P.n[0] / V;                     // vectors along the fastest free mode (stride 1 in A and C)
  const uint64_t n0p = P.n0p;                          // lanes per row: n0v rounded up to whole warps
  const uint64_t total_p = (P.total / P.n[0]) * n0p;

  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total_p; j += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t rem = j / n0p;
    const uint64_t i0 = j - rem * n0p;
    if (i0 >= n0v) continue;                           // surplus lane of the row's last warp
    uint64_t offa = i0 * V, offc = offa;
    for (uint32_t d = 1; d < P.nfree; ++d) {
      const uint64_t i = rem % P.n[d];
      rem /= P.n[d];
      offa += i * P.wa[d];
      offc += i * P.wc[d];
    }
    const T* ap = A + offa;
    T acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = Num<T>::zero();
    uint64_t k = 0;
    for (; k + KU <= P.nq; k += KU) {
      Vec<T, V> v[KU];
#pragma unroll
      for (int s = 0; s < KU; ++s) v[s] = load_stream<T, V>(ap + (k + s) * P.wq);
#pragma unroll
      for (int s = 0; s < KU; ++s) {
        const T bb = B[k + s];
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] = Num<T>::madd(v[s].e[e], bb, acc[e]);
      }
    }
    if (k < P.nq) {                                     // last, partial batch: still all loads first
      Vec<T, V> v[KU];
#pragma unroll
      for (int s = 0; s < KU; ++s)
        if (k + s < P.nq) v[s] = load_stream<T, V>(ap + (k + s) * P.wq);
#pragma unroll
      for (int s = 0; s < KU; ++s)
        if (k + s < P.nq) {
          const T bb = B[k + s];
#pragma unroll
          for (int e = 0; e < V; ++e) acc[e] = Num<T>::madd(v[s].e[e], bb, acc[e]);
        }
    }
    Vec<T, V>* out = reinterpret_cast<Vec<T, V>*>(C + offc);
    Vec<T, V> r;
    if (P.accumulate) {
      r = *out;
#pragma unroll
      for (int e = 0; e < V; ++e) r.e[e] = Num<T>::add(r.e[e], acc[e]);
    } else {
#pragma unroll
      for (int e = 0; e < V; ++e) r.e[e] = acc[e];
    }
    *out = r;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// STRIDED, q contiguous (wa[q-1] == 1): the fibers are contiguous runs of n_q elements whose starts are strided
// (rows of a matrix with a padded leading dimension, the fastest mode of a sliced tensor).  One thread per output
// would put the lanes of a warp on 32 different fibers (measured 256^3 x 250 slice, q = 1: 1.1 TB/s).  Here a GROUP of
// G lanes (power of two <= 32) shares a fiber: the lanes read consecutive vectors of V elements, 64 bytes of loads in flight
// each, and combine their partial sums with a __shfl_xor butterfly; a warp works on 32 / G fibers at a time.
// V > 1 requires n_q, the strides of the free modes and the addresses of A and b to be multiples of V.
// ------------------------------------------------------------------------------------------------------------------
template<class T, int V> __host__ __device__ constexpr int strided_dot_ku() { return sizeof(T) * V >= 16 ? 4 : sizeof(T) * V >= 8 ? 8 : 16; }

template<class T, int V>
__global__ void __launch_bounds__(256)
ttv_strided_dot_kernel(const StridedParams P, const uint32_t G)
{
  pdl_prologue();
  const T* __restrict__ A = static_cast<const T*>(P.a);
  const T* __restrict__ B = static_cast<const T*>(P.b);
  T* __restrict__       C = static_cast<T*>(P.c);
  constexpr int KU = strided_dot_ku<T, V>();                  // 64 bytes of loads in flight per lane

  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t g    = lane % G;                             // position inside the fiber's lane group
  const uint32_t fpw  = 32u / G;                              // fibers a warp works on at a time
  const uint64_t kv   = P.nq / V;                             // vectors per fiber
  const uint64_t warp0 = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);

  for (uint64_t base = warp0 * fpw; base < P.total; base += warps * fpw) {       // warp-uniform trip count
    const uint64_t j = base + lane / G;
    const bool valid = j < P.total;
    uint64_t offa, offc;
    strided_decode(P, valid ? j : 0, offa, offc);
    const T* ap = A + offa;
    T acc = Num<T>::zero();
    for (uint64_t k = g; k < kv; k += (uint64_t)G * KU) {
      Vec<T, V> v[KU], bv[KU];
#pragma unroll
      for (int s = 0; s < KU; ++s) {
        const uint64_t kk = k + (uint64_t)s * G;
        if (valid && kk < kv) {
          v[s]  = *reinterpret_cast<const Vec<T, V>*>(ap + kk * V);
          bv[s] = *reinterpret_cast<const Vec<T, V>*>(B + kk * V);
        } else {
#pragma unroll
          for (int e = 0; e < V; ++e) { v[s].e[e] = Num<T>::zero(); bv[s].e[e] = Num<T>::zero(); }
        }
      }
#pragma unroll
      for (int s = 0; s < KU; ++s)
#pragma unroll
        for (int e = 0; e < V; ++e) acc = Num<T>::madd(v[s].e[e], bv[s].e[e], acc);
    }
    for (uint32_t h = G >> 1; h > 0; h >>= 1) {
      // butterfly inside the group (xor masks below G never leave it)
      static_assert(sizeof(T) % 4 == 0, "element size");
      uint32_t w[sizeof(T) / 4];
      memcpy(w, &acc, sizeof(T));
      T other;
#pragma unroll
      for (unsigned i = 0; i < sizeof(T) / 4; ++i) w[i] = __shfl_xor_sync(0xffffffffu, w[i], (int)h);
      memcpy(&other, w, sizeof(T));
      acc = Num<T>::add(acc, other);
    }
    if (valid && g == 0) {
      T* out = C + offc;
      *out = P.accumulate ? Num<T>::add(*out, acc) : acc;
    }
  }
}

} // namespace ttvb
