// kernels.cuh -- the sm_100a TTV kernels on the canonical view  A[outer][nq][inner], b[nq], C[outer][inner].
//
// They replace the reference's arithmetic layer (include/tlib/detail/matrix_times_vector.h: gemv_row :51-91,
// gemv_col :108-179, dot :264-295, the BLAS calls :213-256) and its OpenMP loop nest
// (detail/tensor_times_vector.h:189-398) in one launch.
//
// Thread tile.  A CTA of `threads` threads is arranged as (to, ty, tx), tx fastest:
//     tx  threads along inner (COL) -- each owns V contiguous outputs, so a warp reads a contiguous run of a row of A
//     ty  threads along n_q         -- thread ty visits k = ty, ty+TY, ...   (DOT: vectors of V consecutive k)
//     to  threads along outer       -- several slabs per CTA when one slab is smaller than the CTA
// When tx*V == inner the lanes (ty, tx) of a warp cover consecutive rows, i.e. one contiguous run of memory, which is
// how small inner extents stay coalesced.  The ty partial sums are combined by warp shuffles (DOT, ty <= 32) or a
// shared-memory tree; n_q partitions across CTAs (ksplit > 1) go to a workspace and are summed by ttv_reduce_kernel
// in fixed order, so results are deterministic.
//
// Traffic: every element of A is loaded exactly once with ld.global.nc.L1::no_allocate (16 bytes when alignment
// allows), b is staged in shared memory once per CTA (hoisted out of the tile loop when it fits), C is written once.
#pragma once

#include "numeric.cuh"

namespace ttvb {

struct TileParams {
  const void* a;
  const void* b;
  void*       c;          // C, or the workspace [ksplit][outer*inner] when ksplit > 1
  uint64_t outer, nq, inner;
  uint64_t kchunk;        // n_q elements per partition
  uint64_t itiles, otiles, tiles;
  uint32_t tx, ty, to;
  uint32_t ksplit;
  uint32_t kb;            // elements of b per shared-memory chunk
  uint32_t accumulate;    // C += (only honoured when ksplit == 1; otherwise the reduce pass does it)
};

// ------------------------------------------------------------------------------------------------------------------
// COL: column GEMV, vector of V outputs along inner per thread.
// ------------------------------------------------------------------------------------------------------------------
template<class T, int V, int KU>
__global__ void __launch_bounds__(256, 4)
ttv_col_kernel(const TileParams P)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sb  = reinterpret_cast<T*>(smem_raw);          // [kb]
  T* red = sb + P.kb;                               // [threads][V]

  const T* __restrict__ A = static_cast<const T*>(P.a);
  const T* __restrict__ B = static_cast<const T*>(P.b);
  T* __restrict__       C = static_cast<T*>(P.c);

  const uint32_t tid = threadIdx.x;
  const uint32_t tx  = tid % P.tx;
  const uint32_t ty  = (tid / P.tx) % P.ty;
  const uint32_t to  = tid / (P.tx * P.ty);
  const bool     live = to < P.to;
  const uint64_t kstride = (uint64_t)P.ty * P.inner;      // elements between two k visited by one thread
  const bool     b_resident = (P.ksplit == 1) && (P.nq <= P.kb);   // b fits: stage it once per CTA

  if (b_resident) {
    for (uint32_t j = tid; j < (uint32_t)P.nq; j += blockDim.x) sb[j] = B[j];
    __syncthreads();
  }

  for (uint64_t tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
    const uint64_t it = tile % P.itiles;
    const uint64_t r  = tile / P.itiles;
    const uint32_t ks = (uint32_t)(r % P.ksplit);
    const uint64_t ot = r / P.ksplit;
    const uint64_t o  = ot * P.to + to;
    const uint64_t i0 = (it * P.tx + tx) * V;
    const bool act = live && o < P.outer && i0 < P.inner;
    const uint64_t kbeg = (uint64_t)ks * P.kchunk;
    const uint64_t kend = min(kbeg + P.kchunk, P.nq);

    T acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = Num<T>::zero();

    for (uint64_t k0 = kbeg; k0 < kend; k0 += P.kb) {
      const uint32_t kn = (uint32_t)min((uint64_t)P.kb, kend - k0);
      if (!b_resident) {
        __syncthreads();
        for (uint32_t j = tid; j < kn; j += blockDim.x) sb[j] = B[k0 + j];
        __syncthreads();
      }
      if (act) {
        const T* ap = A + (o * P.nq + k0 + ty) * P.inner + i0;
        uint32_t k = ty;
        // main loop: KU independent vector loads in flight, then KU*V multiply-adds
        for (; k + (KU - 1) * P.ty < kn; k += KU * P.ty, ap += KU * kstride) {
          Vec<T, V> v[KU];
#pragma unroll
          for (int u = 0; u < KU; ++u) v[u] = load_stream<T, V>(ap + u * kstride);
#pragma unroll
          for (int u = 0; u < KU; ++u) {
            const T bb = sb[k + u * P.ty];
#pragma unroll
            for (int j = 0; j < V; ++j) acc[j] = Num<T>::madd(v[u].e[j], bb, acc[j]);
          }
        }
        for (; k < kn; k += P.ty, ap += kstride) {
          const Vec<T, V> v = load_stream<T, V>(ap);
          const T bb = sb[k];
#pragma unroll
          for (int j = 0; j < V; ++j) acc[j] = Num<T>::madd(v.e[j], bb, acc[j]);
        }
      }
    }

    // combine the ty partial sums of each output: shared-memory tree over ty
    if (P.ty > 1) {
      __syncthreads();
#pragma unroll
      for (int j = 0; j < V; ++j) red[tid * V + j] = acc[j];
      __syncthreads();
      uint32_t span = 1;
      while (span < P.ty) span <<= 1;
      for (uint32_t h = span >> 1; h > 0; h >>= 1) {
        if (live && ty < h && ty + h < P.ty) {
#pragma unroll
          for (int j = 0; j < V; ++j)
            red[tid * V + j] = Num<T>::add(red[tid * V + j], red[(tid + h * P.tx) * V + j]);
        }
        __syncthreads();
      }
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = red[tid * V + j];
    }

    if (act && ty == 0) {
      T* dst = C + (P.ksplit > 1 ? (uint64_t)ks * P.outer * P.inner : 0) + o * P.inner + i0;
      Vec<T, V> outv;
      if (P.accumulate && P.ksplit == 1) {
        const Vec<T, V> old = *reinterpret_cast<const Vec<T, V>*>(dst);
#pragma unroll
        for (int j = 0; j < V; ++j) outv.e[j] = Num<T>::add(old.e[j], acc[j]);
      } else {
#pragma unroll
        for (int j = 0; j < V; ++j) outv.e[j] = acc[j];
      }
      *reinterpret_cast<Vec<T, V>*>(dst) = outv;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// DOT: mode q is the contiguous one (inner == 1).  ty lanes cooperate on one fiber with vectors of V consecutive k;
// `to` fibers per CTA.  With ty <= 32 a fiber lives inside one warp and the partial sums are combined with
// __shfl_xor_sync; larger ty uses the shared-memory tree.
// ------------------------------------------------------------------------------------------------------------------
template<class T>
__device__ __forceinline__ T shfl_xor_elem(T v, int mask)
{
  static_assert(sizeof(T) % 4 == 0, "element size");
  uint32_t w[sizeof(T) / 4];
  memcpy(w, &v, sizeof(T));
#pragma unroll
  for (unsigned i = 0; i < sizeof(T) / 4; ++i) w[i] = __shfl_xor_sync(0xffffffffu, w[i], mask);
  T r;
  memcpy(&r, w, sizeof(T));
  return r;
}

template<class T, int V, int KU>
__global__ void __launch_bounds__(256, 4)
ttv_dot_kernel(const TileParams P)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sb  = reinterpret_cast<T*>(smem_raw);          // [kb]
  T* red = sb + P.kb;                               // [threads]

  const T* __restrict__ A = static_cast<const T*>(P.a);
  const T* __restrict__ B = static_cast<const T*>(P.b);
  T* __restrict__       C = static_cast<T*>(P.c);

  const uint32_t tid = threadIdx.x;
  const uint32_t ty  = tid % P.ty;
  const uint32_t to  = tid / P.ty;
  const bool     live = to < P.to;
  const uint32_t kstep = P.ty * V;                   // n_q elements one pass of the fiber's lanes covers
  const bool     b_resident = (P.ksplit == 1) && (P.nq <= P.kb);

  if (b_resident) {
    for (uint32_t j = tid; j < (uint32_t)P.nq; j += blockDim.x) sb[j] = B[j];
    __syncthreads();
  }

  for (uint64_t tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
    const uint32_t ks = (uint32_t)(tile % P.ksplit);
    const uint64_t ot = tile / P.ksplit;
    const uint64_t o  = ot * P.to + to;
    const bool act = live && o < P.outer;
    const uint64_t kbeg = (uint64_t)ks * P.kchunk;
    const uint64_t kend = min(kbeg + P.kchunk, P.nq);

    T acc = Num<T>::zero();

    for (uint64_t k0 = kbeg; k0 < kend; k0 += P.kb) {
      const uint32_t kn = (uint32_t)min((uint64_t)P.kb, kend - k0);   // multiple of V (nq % V == 0, kb % V == 0)
      if (!b_resident) {
        __syncthreads();
        for (uint32_t j = tid; j < kn; j += blockDim.x) sb[j] = B[k0 + j];
        __syncthreads();
      }
      if (act) {
        const T* ap = A + o * P.nq + k0 + (uint64_t)ty * V;
        uint32_t k = ty * V;
        for (; k + (KU - 1) * kstep < kn; k += KU * kstep, ap += KU * kstep) {
          Vec<T, V> v[KU];
#pragma unroll
          for (int u = 0; u < KU; ++u) v[u] = load_stream<T, V>(ap + u * kstep);
#pragma unroll
          for (int u = 0; u < KU; ++u) {
            const Vec<T, V> bv = *reinterpret_cast<const Vec<T, V>*>(sb + k + u * kstep);
#pragma unroll
            for (int j = 0; j < V; ++j) acc = Num<T>::madd(v[u].e[j], bv.e[j], acc);
          }
        }
        for (; k < kn; k += kstep, ap += kstep) {
          const Vec<T, V> v  = load_stream<T, V>(ap);
          const Vec<T, V> bv = *reinterpret_cast<const Vec<T, V>*>(sb + k);
#pragma unroll
          for (int j = 0; j < V; ++j) acc = Num<T>::madd(v.e[j], bv.e[j], acc);
        }
      }
    }

    if (P.ty > 1 && P.ty <= 32) {
      // ty is a power of two <= 32 and divides the warp: butterfly inside the fiber's lane group
      for (uint32_t h = P.ty >> 1; h > 0; h >>= 1) acc = Num<T>::add(acc, shfl_xor_elem(acc, (int)h));
    } else if (P.ty > 32) {
      __syncthreads();
      red[tid] = acc;
      __syncthreads();
      for (uint32_t h = P.ty >> 1; h > 0; h >>= 1) {     // ty is a power of two
        if (live && ty < h) red[tid] = Num<T>::add(red[tid], red[tid + h]);
        __syncthreads();
      }
      acc = red[tid];
    }

    if (act && ty == 0) {
      T* dst = C + (P.ksplit > 1 ? (uint64_t)ks * P.outer : 0) + o;
      *dst = (P.accumulate && P.ksplit == 1) ? Num<T>::add(*dst, acc) : acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// second pass of the split-n_q variant: C[j] (+)= sum_s ws[s][j], s in fixed order
// ------------------------------------------------------------------------------------------------------------------
template<class T>
__global__ void __launch_bounds__(256)
ttv_reduce_kernel(const T* __restrict__ ws, T* __restrict__ c, uint64_t n, uint32_t ksplit, uint32_t accumulate)
{
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) {
    T s = accumulate ? c[j] : Num<T>::zero();
    for (uint32_t p = 0; p < ksplit; ++p) s = Num<T>::add(s, ws[(uint64_t)p * n + j]);
    c[j] = s;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// synthetic data on the device (same generator as oracle/ttv_oracle.c:ttv_oracle_fill)
// ------------------------------------------------------------------------------------------------------------------
template<class T> __device__ __forceinline__ T synth(uint64_t seed, uint64_t j);
template<> __device__ __forceinline__ float    synth<float>(uint64_t seed, uint64_t j)    { return (float)unit_pm1(splitmix64(seed ^ j)); }
template<> __device__ __forceinline__ double   synth<double>(uint64_t seed, uint64_t j)   { return unit_pm1(splitmix64(seed ^ j)); }
template<> __device__ __forceinline__ cf32     synth<cf32>(uint64_t seed, uint64_t j)     { return cf32{(float)unit_pm1(splitmix64(seed ^ (2 * j))), (float)unit_pm1(splitmix64(seed ^ (2 * j + 1)))}; }
template<> __device__ __forceinline__ cf64     synth<cf64>(uint64_t seed, uint64_t j)     { return cf64{unit_pm1(splitmix64(seed ^ (2 * j))), unit_pm1(splitmix64(seed ^ (2 * j + 1)))}; }
template<> __device__ __forceinline__ uint32_t synth<uint32_t>(uint64_t seed, uint64_t j) { return (uint32_t)((int32_t)(splitmix64(seed ^ j) % 17u) - 8); }
template<> __device__ __forceinline__ unsigned long long synth<unsigned long long>(uint64_t seed, uint64_t j) { return (unsigned long long)((long long)(splitmix64(seed ^ j) % 17u) - 8); }

template<class T>
__global__ void __launch_bounds__(256)
ttv_fill_kernel(T* __restrict__ x, uint64_t first, uint64_t count, uint64_t seed)
{
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x)
    x[i] = synth<T>(seed, first + i);
}

} // namespace ttvb
