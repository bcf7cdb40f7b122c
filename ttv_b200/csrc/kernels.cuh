// kernels.cuh -- the sm_100a TTV kernels on the canonical view  A[outer][nq][inner], b[nq], C[outer][inner].
//
// They replace the reference's arithmetic layer (include/tlib/detail/matrix_times_vector.h: gemv_row :51-91,
// gemv_col :108-179, dot :264-295, the BLAS calls :213-256) and its OpenMP loop nest
// (detail/tensor_times_vector.h:189-398) in one launch.
//
// Thread tile.  A CTA of `threads` threads is arranged as (to, ty, tx), tx fastest:
//     tx  threads along inner (COL) -- each owns V contiguous outputs, so a warp reads a contiguous run of a row of A
//     ty  threads along n_q         -- thread ty visits k = ty, ty+TY, ...   (DOT: vectors of V consecutive k)
//     to  threads along outer       -- several slabs / fibers per CTA when one of them is smaller than the CTA
// When tx*V == inner the lanes (ty, tx) of a warp cover consecutive rows, i.e. one contiguous run of memory, which is
// how small inner extents stay coalesced.
//
// Loads in flight.  The path is HBM-bound, so what matters is bytes in flight per SM (Little's law: ~45 KB per SM at
// 7+ TB/s).  Every thread issues one BATCH of NU x KU independent loads (128 bytes with 16-byte vectors) before it
// consumes any of them: KU steps along n_q for each of NU independent outputs ("units": further tiles along inner, or
// further slabs / fibers along outer).  Short contractions (n_q = 2, 4, ...) therefore run with more units instead of
// starving.  Full batches take an unpredicated fast path; the edges (last batch, last tile) are predicated -- there is
// no scalar remainder loop.
//
// Reductions.  The ty partial sums of an output are combined by warp shuffles (DOT, ty <= 32) or a shared-memory
// tree; n_q partitions across CTAs (ksplit > 1) go to a workspace and are summed by ttv_reduce_kernel in fixed order,
// so results are deterministic.
//
// Traffic: every element of A is loaded exactly once, b is staged in shared memory once per CTA (hoisted out of the
// tile loop when it fits), C is written once.
#pragma once

#include "numeric.cuh"

namespace ttvb {

// CTAs of 256 threads per SM the register budget is planned for: 4 -> 64 registers, 3 -> 80, 2 -> 128.
#ifdef TTVB_MIN_CTAS
template<int NU, int KU> constexpr int min_ctas() { return TTVB_MIN_CTAS; }
#else
template<int NU, int KU> constexpr int min_ctas() { return 3; }
#endif

// the column kernel also exists with 256 bytes in flight per thread (16 vector loads): 2 CTAs per SM, 128 registers
template<class T, int V, int NU, int KU> constexpr int col_min_ctas() { return NU * KU * V * (int)sizeof(T) > 128 ? 2 : min_ctas<NU, KU>(); }

struct TileParams {
  const void* a;
  const void* b;
  void*       c;          // C, or the workspace [ksplit][outer*inner] when ksplit > 1
  uint64_t outer, nq, inner;
  uint64_t kchunk;        // n_q elements per partition
  uint64_t itiles, otiles, tiles;
  uint64_t a_ustride;     // elements between two units of one thread in A
  uint64_t c_ustride;     // ... and in C  (COLX: the output columns owned by one tile)
  uint32_t tx, ty, to;
  uint32_t ksplit;
  uint32_t kb;            // elements of b per shared-memory chunk
  uint32_t accumulate;    // C += (only honoured when ksplit == 1; otherwise the reduce pass does it)
  uint32_t udir;          // units run along inner (0) or along outer (1)
  uint32_t stream;        // 1: L1::no_allocate loads
};

template<class T, int V>
__device__ __forceinline__ Vec<T, V> load_a(const T* p, bool stream)
{
  Vec<T, V> v;
  if (stream) Ld<sizeof(T) * V>::nc(&v, p);
  else        v = *reinterpret_cast<const Vec<T, V>*>(p);      // read-only path through L1 (A is const __restrict__)
  return v;
}

template<class T, int V>
__device__ __forceinline__ Vec<T, V> zero_vec()
{
  Vec<T, V> v;
#pragma unroll
  for (int j = 0; j < V; ++j) v.e[j] = Num<T>::zero();
  return v;
}

// ------------------------------------------------------------------------------------------------------------------
// COL: column GEMV, vector of V outputs along inner per thread and unit; NU units x KU k-steps in flight.
// ------------------------------------------------------------------------------------------------------------------
template<class T, int V, int NU, int KU, bool PRED>
__device__ __forceinline__ void col_batch(T (&acc)[NU][V], const T* ap, uint64_t a_ustride, uint64_t kstride, const T* sb,
                                          uint32_t k, uint32_t tyn, uint32_t kn, int nvalid, bool stream)
{
  Vec<T, V> v[NU][KU];
#pragma unroll
  for (int u = 0; u < NU; ++u)
#pragma unroll
    for (int s = 0; s < KU; ++s) {
      if constexpr (PRED)
        v[u][s] = (u < nvalid && k + s * tyn < kn) ? load_a<T, V>(ap + u * a_ustride + s * kstride, stream) : zero_vec<T, V>();
      else
        v[u][s] = load_a<T, V>(ap + u * a_ustride + s * kstride, stream);
    }
#pragma unroll
  for (int s = 0; s < KU; ++s) {
    T bb;
    if constexpr (PRED) bb = (k + s * tyn < kn) ? sb[k + s * tyn] : Num<T>::zero();
    else bb = sb[k + s * tyn];
#pragma unroll
    for (int u = 0; u < NU; ++u)
#pragma unroll
      for (int j = 0; j < V; ++j) acc[u][j] = Num<T>::madd(v[u][s].e[j], bb, acc[u][j]);
  }
}

// The same batch with b taken straight from global memory / L2 (bg = B + k0): lanes strung along n_q read consecutive
// elements of b, and these loads are issued together with those of A.
template<class T, int V, int NU, int KU, bool PRED>
__device__ __forceinline__ void col_batch_bg(T (&acc)[NU][V], const T* ap, uint64_t a_ustride, uint64_t kstride, const T* bg,
                                             uint32_t k, uint32_t tyn, uint32_t kn, int nvalid, bool stream)
{
  Vec<T, V> v[NU][KU];
  T bb[KU];
#pragma unroll
  for (int s = 0; s < KU; ++s) {
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      if constexpr (PRED)
        v[u][s] = (u < nvalid && k + s * tyn < kn) ? load_a<T, V>(ap + u * a_ustride + s * kstride, stream) : zero_vec<T, V>();
      else
        v[u][s] = load_a<T, V>(ap + u * a_ustride + s * kstride, stream);
    }
    if constexpr (PRED) bb[s] = (k + s * tyn < kn) ? bg[k + s * tyn] : Num<T>::zero();
    else bb[s] = bg[k + s * tyn];
  }
#pragma unroll
  for (int s = 0; s < KU; ++s)
#pragma unroll
    for (int u = 0; u < NU; ++u)
#pragma unroll
      for (int j = 0; j < V; ++j) acc[u][j] = Num<T>::madd(v[u][s].e[j], bb[s], acc[u][j]);
}

template<class T, int V, int NU, int KU, bool BG = false>
__global__ void __launch_bounds__(256, col_min_ctas<T, V, NU, KU>())
ttv_col_kernel(const TileParams P)
{
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sb  = reinterpret_cast<T*>(smem_raw);          // [kb]   (nothing when b is read directly)
  T* red = sb + (BG ? 0u : P.kb);                   // [threads][NU*V]

  const T* __restrict__ A = static_cast<const T*>(P.a);
  const T* __restrict__ B = static_cast<const T*>(P.b);
  T* __restrict__       C = static_cast<T*>(P.c);

  const uint32_t tid = threadIdx.x;
  const uint32_t tx  = tid % P.tx;
  const uint32_t ty  = (tid / P.tx) % P.ty;
  const uint32_t to  = tid / (P.tx * P.ty);
  const bool     live = to < P.to;
  const bool     stream = P.stream != 0;
  const uint64_t kstride = (uint64_t)P.ty * P.inner;      // elements between two k visited by one thread
  constexpr bool bdirect = BG;                                     // b is not staged: kb spans the whole partition
  const bool     b_resident = !bdirect && (P.ksplit == 1) && (P.nq <= P.kb);   // b fits: stage it once per CTA

  if (b_resident) {
    for (uint32_t j = tid; j < (uint32_t)P.nq; j += blockDim.x) sb[j] = B[j];
    __syncthreads();
  }

  for (uint64_t tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
    const uint64_t it = tile % P.itiles;
    const uint64_t r  = tile / P.itiles;
    const uint32_t ks = (uint32_t)(r % P.ksplit);
    const uint64_t ot = r / P.ksplit;
    // first unit of this thread; unit u adds u*tx*V along inner (udir 0) or u*to along outer (udir 1)
    const uint64_t o  = (P.udir ? ot * NU : ot) * P.to + to;
    const uint64_t i0 = ((P.udir ? it : it * NU) * P.tx + tx) * V;
    int nvalid = 0;                                   // units [0, nvalid) of this thread exist
    if (live && o < P.outer && i0 < P.inner) {
      const uint64_t room = P.udir ? (P.outer - o + P.to - 1) / P.to
                                   : (P.inner - i0 + (uint64_t)P.tx * V - 1) / ((uint64_t)P.tx * V);
      nvalid = room < (uint64_t)NU ? (int)room : NU;
    }
    const uint64_t kbeg = (uint64_t)ks * P.kchunk;
    const uint64_t kend = min(kbeg + P.kchunk, P.nq);

    T acc[NU][V];
#pragma unroll
    for (int u = 0; u < NU; ++u)
#pragma unroll
      for (int j = 0; j < V; ++j) acc[u][j] = Num<T>::zero();

    for (uint64_t k0 = kbeg; k0 < kend; k0 += P.kb) {
      const uint32_t kn = (uint32_t)min((uint64_t)P.kb, kend - k0);
      if (!b_resident && !bdirect) {       // (b_resident is false when bdirect)
        __syncthreads();
        for (uint32_t j = tid; j < kn; j += blockDim.x) sb[j] = B[k0 + j];
        __syncthreads();
      }
      if (nvalid > 0) {
        const T* ap = A + (o * P.nq + k0 + ty) * P.inner + i0;
        uint32_t k = ty;
        if constexpr (bdirect) {
          const T* bg = B + k0;
          if (nvalid == NU)
            for (; k + (KU - 1) * P.ty < kn; k += KU * P.ty, ap += KU * kstride)
              col_batch_bg<T, V, NU, KU, false>(acc, ap, P.a_ustride, kstride, bg, k, P.ty, kn, nvalid, stream);
          for (; k < kn; k += KU * P.ty, ap += KU * kstride)
            col_batch_bg<T, V, NU, KU, true>(acc, ap, P.a_ustride, kstride, bg, k, P.ty, kn, nvalid, stream);
        } else {
          if (nvalid == NU)
            for (; k + (KU - 1) * P.ty < kn; k += KU * P.ty, ap += KU * kstride)          // full batches
              col_batch<T, V, NU, KU, false>(acc, ap, P.a_ustride, kstride, sb, k, P.ty, kn, nvalid, stream);
          for (; k < kn; k += KU * P.ty, ap += KU * kstride)                              // edges
            col_batch<T, V, NU, KU, true>(acc, ap, P.a_ustride, kstride, sb, k, P.ty, kn, nvalid, stream);
        }
      }
    }

    // combine the ty partial sums of each output: shared-memory tree over ty
    if (P.ty > 1) {
      T* mine = red + (size_t)tid * (NU * V);
      __syncthreads();
#pragma unroll
      for (int u = 0; u < NU; ++u)
#pragma unroll
        for (int j = 0; j < V; ++j) mine[u * V + j] = acc[u][j];
      __syncthreads();
      uint32_t span = 1;
      while (span < P.ty) span <<= 1;
      for (uint32_t h = span >> 1; h > 0; h >>= 1) {
        if (live && ty < h && ty + h < P.ty) {
          const T* other = mine + (size_t)h * P.tx * (NU * V);
#pragma unroll
          for (int e = 0; e < NU * V; ++e) mine[e] = Num<T>::add(mine[e], other[e]);
        }
        __syncthreads();
      }
#pragma unroll
      for (int u = 0; u < NU; ++u)
#pragma unroll
        for (int j = 0; j < V; ++j) acc[u][j] = mine[u * V + j];
    }

    if (ty == 0) {
      T* dst = C + (P.ksplit > 1 ? (uint64_t)ks * P.outer * P.inner : 0) + o * P.inner + i0;
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        if (u < nvalid) {
          Vec<T, V>* out = reinterpret_cast<Vec<T, V>*>(dst + u * P.c_ustride);
          Vec<T, V> val;
          if (P.accumulate && P.ksplit == 1) {
            const Vec<T, V> old = *out;
#pragma unroll
            for (int j = 0; j < V; ++j) val.e[j] = Num<T>::add(old.e[j], acc[u][j]);
          } else {
#pragma unroll
            for (int j = 0; j < V; ++j) val.e[j] = acc[u][j];
          }
          *out = val;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// DOT: mode q is the contiguous one (inner == 1).  ty lanes cooperate on one fiber with vectors of V consecutive k;
// a CTA holds `to` lane groups and every group works on NU fibers at once (NU*KU loads in flight per lane).  With
// ty <= 32 a fiber lives inside one warp and the partial sums are combined with __shfl_xor_sync; larger ty (few, very
// long fibers) uses the shared-memory tree.
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

template<class T>
__device__ __forceinline__ T shfl_down_elem(T v, int delta)
{
  static_assert(sizeof(T) % 4 == 0, "element size");
  uint32_t w[sizeof(T) / 4];
  memcpy(w, &v, sizeof(T));
#pragma unroll
  for (unsigned i = 0; i < sizeof(T) / 4; ++i) w[i] = __shfl_down_sync(0xffffffffu, w[i], delta);
  T r;
  memcpy(&r, w, sizeof(T));
  return r;
}

// combines the ty partial sums of NU fibers and stores them (shared by the two DOT kernels)
template<class T, int NU>
__device__ __forceinline__ void dot_finish(T (&acc)[NU], const TileParams& P, T* red, T* C, uint32_t tid, uint32_t ty, bool live,
                                           uint64_t o, uint32_t ks, int nvalid)
{
  if (P.ty > 1 && P.ty <= 32) {
    // ty is a power of two <= 32 and divides the warp: butterfly inside the fiber's lane group
    for (uint32_t h = P.ty >> 1; h > 0; h >>= 1) {
#pragma unroll
      for (int u = 0; u < NU; ++u) acc[u] = Num<T>::add(acc[u], shfl_xor_elem(acc[u], (int)h));
    }
  } else if (P.ty > 32) {
    T* mine = red + (size_t)tid * NU;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < NU; ++u) mine[u] = acc[u];
    __syncthreads();
    for (uint32_t h = P.ty >> 1; h > 0; h >>= 1) {     // ty is a power of two
      if (live && ty < h) {
#pragma unroll
        for (int u = 0; u < NU; ++u) mine[u] = Num<T>::add(mine[u], mine[(size_t)h * NU + u]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < NU; ++u) acc[u] = mine[u];
  }
  if (ty == 0) {
    T* dst = C + (P.ksplit > 1 ? (uint64_t)ks * P.outer : 0) + o;
#pragma unroll
    for (int u = 0; u < NU; ++u)
      if (u < nvalid) {
        T* out = dst + (uint64_t)u * P.to;
        *out = (P.accumulate && P.ksplit == 1) ? Num<T>::add(*out, acc[u]) : acc[u];
      }
  }
}

template<class T, int V, int NU, int KU, bool PRED>
__device__ __forceinline__ void dot_batch(T (&acc)[NU], const T* ap, uint64_t a_ustride, uint32_t kstep, const T* sb,
                                          uint32_t k, uint32_t kn, int nvalid, bool stream)
{
  Vec<T, V> v[NU][KU];
#pragma unroll
  for (int u = 0; u < NU; ++u)
#pragma unroll
    for (int s = 0; s < KU; ++s) {
      if constexpr (PRED)
        v[u][s] = (u < nvalid && k + s * kstep < kn) ? load_a<T, V>(ap + u * a_ustride + s * kstep, stream) : zero_vec<T, V>();
      else
        v[u][s] = load_a<T, V>(ap + u * a_ustride + s * kstep, stream);
    }
#pragma unroll
  for (int s = 0; s < KU; ++s) {
    Vec<T, V> bv;
    if constexpr (PRED) bv = (k + s * kstep < kn) ? *reinterpret_cast<const Vec<T, V>*>(sb + k + s * kstep) : zero_vec<T, V>();
    else bv = *reinterpret_cast<const Vec<T, V>*>(sb + k + s * kstep);
#pragma unroll
    for (int u = 0; u < NU; ++u)
#pragma unroll
      for (int j = 0; j < V; ++j) acc[u] = Num<T>::madd(v[u][s].e[j], bv.e[j], acc[u]);
  }
}

// b vectors straight from global memory / L2 (bg = B + k0), issued together with the loads of A
template<class T, int V, int NU, int KU, bool PRED>
__device__ __forceinline__ void dot_batch_bg(T (&acc)[NU], const T* ap, uint64_t a_ustride, uint32_t kstep, const T* bg,
                                             uint32_t k, uint32_t kn, int nvalid, bool stream)
{
  Vec<T, V> v[NU][KU];
  Vec<T, V> bv[KU];
#pragma unroll
  for (int s = 0; s < KU; ++s) {
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      if constexpr (PRED)
        v[u][s] = (u < nvalid && k + s * kstep < kn) ? load_a<T, V>(ap + u * a_ustride + s * kstep, stream) : zero_vec<T, V>();
      else
        v[u][s] = load_a<T, V>(ap + u * a_ustride + s * kstep, stream);
    }
    if constexpr (PRED) bv[s] = (k + s * kstep < kn) ? *reinterpret_cast<const Vec<T, V>*>(bg + k + s * kstep) : zero_vec<T, V>();
    else bv[s] = *reinterpret_cast<const Vec<T, V>*>(bg + k + s * kstep);
  }
#pragma unroll
  for (int s = 0; s < KU; ++s)
#pragma unroll
    for (int u = 0; u < NU; ++u)
#pragma unroll
      for (int j = 0; j < V; ++j) acc[u] = Num<T>::madd(v[u][s].e[j], bv[s].e[j], acc[u]);
}

template<class T, int V, int NU, int KU, bool BG = false>
__global__ void __launch_bounds__(256, col_min_ctas<T, V, NU, KU>())
ttv_dot_kernel(const TileParams P)
{
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sb  = reinterpret_cast<T*>(smem_raw);          // [kb]   (nothing when b is read directly)
  T* red = sb + (BG ? 0u : P.kb);                   // [threads][NU]

  const T* __restrict__ A = static_cast<const T*>(P.a);
  const T* __restrict__ B = static_cast<const T*>(P.b);
  T* __restrict__       C = static_cast<T*>(P.c);

  const uint32_t tid = threadIdx.x;
  const uint32_t ty  = tid % P.ty;
  const uint32_t to  = tid / P.ty;
  const bool     live = to < P.to;
  const bool     stream = P.stream != 0;
  const uint32_t kstep = P.ty * V;                   // n_q elements one pass of the fiber's lanes covers
  constexpr bool bdirect = BG;
  const bool     b_resident = !bdirect && (P.ksplit == 1) && (P.nq <= P.kb);

  if (b_resident) {
    for (uint32_t j = tid; j < (uint32_t)P.nq; j += blockDim.x) sb[j] = B[j];
    __syncthreads();
  }

  for (uint64_t tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
    const uint32_t ks = (uint32_t)(tile % P.ksplit);
    const uint64_t ot = tile / P.ksplit;
    const uint64_t o  = ot * NU * P.to + to;         // fiber of unit 0; unit u is fiber o + u*to
    int nvalid = 0;
    if (live && o < P.outer) {
      const uint64_t room = (P.outer - o + P.to - 1) / P.to;
      nvalid = room < (uint64_t)NU ? (int)room : NU;
    }
    const uint64_t kbeg = (uint64_t)ks * P.kchunk;
    const uint64_t kend = min(kbeg + P.kchunk, P.nq);

    T acc[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) acc[u] = Num<T>::zero();

    for (uint64_t k0 = kbeg; k0 < kend; k0 += P.kb) {
      const uint32_t kn = (uint32_t)min((uint64_t)P.kb, kend - k0);   // multiple of V (nq % V == 0, kb % V == 0)
      if (!b_resident && !bdirect) {       // (b_resident is false when bdirect)
        __syncthreads();
        for (uint32_t j = tid; j < kn; j += blockDim.x) sb[j] = B[k0 + j];
        __syncthreads();
      }
      if (nvalid > 0) {
        const T* ap = A + o * P.nq + k0 + (uint64_t)ty * V;
        uint32_t k = ty * V;
        if constexpr (bdirect) {
          const T* bg = B + k0;
          if (nvalid == NU)
            for (; k + (KU - 1) * kstep < kn; k += KU * kstep, ap += KU * kstep)
              dot_batch_bg<T, V, NU, KU, false>(acc, ap, P.a_ustride, kstep, bg, k, kn, nvalid, stream);
          for (; k < kn; k += KU * kstep, ap += KU * kstep)
            dot_batch_bg<T, V, NU, KU, true>(acc, ap, P.a_ustride, kstep, bg, k, kn, nvalid, stream);
        } else {
          if (nvalid == NU)
            for (; k + (KU - 1) * kstep < kn; k += KU * kstep, ap += KU * kstep)
              dot_batch<T, V, NU, KU, false>(acc, ap, P.a_ustride, kstep, sb, k, kn, nvalid, stream);
          for (; k < kn; k += KU * kstep, ap += KU * kstep)
            dot_batch<T, V, NU, KU, true>(acc, ap, P.a_ustride, kstep, sb, k, kn, nvalid, stream);
        }
      }
    }
    dot_finish<T, NU>(acc, P, red, C, tid, ty, live, o, ks, nvalid);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// DOT with peeling: n_q is not a multiple of the vector width, so fibers start at arbitrary offsets inside a 16-byte
// line.  Because consecutive fibers are contiguous, every fiber still consists of [head | aligned 16-byte vectors |
// tail] with head, tail < V elements.  The aligned body is loaded with full vectors, head and tail with scalar loads
// issued in the same batch.  b is kept in shared memory V times, copy h shifted by h elements, so that the b-vector
// matching any aligned body vector is itself an aligned 16-byte shared-memory load.
// Requires ksplit == 1, the whole b resident (V copies), A 16-byte aligned, ty >= 2V-2 lanes or two rounds of peel.
// ------------------------------------------------------------------------------------------------------------------
template<class T, int V, int NU, int KU>
__global__ void __launch_bounds__(256, min_ctas<NU, KU>())
ttv_dot_peel_kernel(const TileParams P)
{
  pdl_prologue();
  static_assert(V > 1, "peeling needs a vector");
  constexpr int HT = 2;                              // head/tail rounds: HT * ty lanes must cover 2V-2 elements
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t nq   = (uint32_t)P.nq;
  const uint32_t bpad = (nq + V - 1) / V * V + V;    // elements per shifted copy (16-byte aligned rows)
  T* sb  = reinterpret_cast<T*>(smem_raw);           // [V][bpad]; copy h holds b[h + i] at i
  T* red = sb + (size_t)V * bpad;                    // [threads][NU]

  const T* __restrict__ A = static_cast<const T*>(P.a);
  const T* __restrict__ B = static_cast<const T*>(P.b);
  T* __restrict__       C = static_cast<T*>(P.c);

  const uint32_t tid = threadIdx.x;
  const uint32_t ty  = tid % P.ty;
  const uint32_t to  = tid / P.ty;
  const bool     live = to < P.to;
  const bool     stream = P.stream != 0;
  const uint32_t vstep = P.ty;                       // body vectors one pass of the fiber's lanes covers

  for (uint32_t j = tid; j < V * bpad; j += blockDim.x) {
    const uint32_t h = j / bpad, i = j - h * bpad;
    sb[j] = (h + i < nq) ? B[h + i] : Num<T>::zero();
  }
  __syncthreads();

  for (uint64_t tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
    const uint64_t o = tile * NU * P.to + to;        // fiber of unit 0; unit u is fiber o + u*to
    int nvalid = 0;
    if (live && o < P.outer) {
      const uint64_t room = (P.outer - o + P.to - 1) / P.to;
      nvalid = room < (uint64_t)NU ? (int)room : NU;
    }

    T acc[NU];
    const T* body[NU];                               // first aligned vector of each fiber
    const T* bsh[NU];                                // the shifted copy of b that matches it
    uint32_t nbody[NU];                              // aligned vectors in each fiber
    T hx[NU][HT], hb[NU][HT];                        // head / tail elements of A and b handled by this lane
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      acc[u] = Num<T>::zero();
      const uint64_t e0 = (o + (uint64_t)u * P.to) * nq;               // first element of the fiber
      const uint32_t head = (uint32_t)((V - (e0 & (V - 1))) & (V - 1));
      const uint32_t nb = u < nvalid ? (nq - head) / V : 0;
      const uint32_t tail = nq - head - nb * V;
      body[u] = A + e0 + head;
      bsh[u] = sb + (size_t)head * bpad;
      nbody[u] = nb;
#pragma unroll
      for (int r = 0; r < HT; ++r) {
        const uint32_t idx = ty + r * P.ty;          // position inside head ++ tail
        const bool on = u < nvalid && idx < head + tail;
        const uint32_t k = idx < head ? idx : nq - tail + (idx - head);
        hx[u][r] = on ? A[e0 + k] : Num<T>::zero();
        hb[u][r] = on ? sb[k] : Num<T>::zero();
      }
    }
    const uint32_t nvec = nq / V;                    // upper bound of nbody[u]
    for (uint32_t j = ty; j < nvec; j += KU * vstep) {
      Vec<T, V> v[NU][KU];
#pragma unroll
      for (int u = 0; u < NU; ++u)
#pragma unroll
        for (int s = 0; s < KU; ++s) {
          if (j + s * vstep < nbody[u]) v[u][s] = load_a<T, V>(body[u] + (size_t)(j + s * vstep) * V, stream);
          else v[u][s] = zero_vec<T, V>();
        }
#pragma unroll
      for (int u = 0; u < NU; ++u)
#pragma unroll
        for (int s = 0; s < KU; ++s) {
          if (j + s * vstep < nbody[u]) {
            const Vec<T, V> bv = *reinterpret_cast<const Vec<T, V>*>(bsh[u] + (size_t)(j + s * vstep) * V);
#pragma unroll
            for (int e = 0; e < V; ++e) acc[u] = Num<T>::madd(v[u][s].e[e], bv.e[e], acc[u]);
          }
        }
    }
#pragma unroll
    for (int u = 0; u < NU; ++u)
#pragma unroll
      for (int r = 0; r < HT; ++r) acc[u] = Num<T>::madd(hx[u][r], hb[u][r], acc[u]);

    dot_finish<T, NU>(acc, P, red, C, tid, ty, live, o, 0u, nvalid);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// second pass of the split-n_q variant: C[j] (+)= sum_s ws[s][j], s in fixed order
// ------------------------------------------------------------------------------------------------------------------
template<class T>
__global__ void __launch_bounds__(256)
ttv_reduce_kernel(const T* __restrict__ ws, T* __restrict__ c, uint64_t n, uint32_t ksplit, uint32_t accumulate, uint64_t stride)
{
  pdl_prologue();
  // ws is [ksplit][stride], the first n entries of every row are summed (stride == n for the split-n_q workspace)
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) {
    T s = accumulate ? c[j] : Num<T>::zero();
    for (uint32_t p = 0; p < ksplit; ++p) s = Num<T>::add(s, ws[(uint64_t)p * stride + j]);
    c[j] = s;
  }
}

// The same pass when there are FEW outputs and MANY partitions (n_q in the hundreds of millions with a tiny inner extent:
// [1, 2^26, 4] has four outputs and 1 184 partitions): one thread per output would walk its partitions alone, ~1 000
// dependent steps, tens of microseconds next to a 150 us product.  Here a CTA owns an output, its threads take the partitions
// p = t, t + 256, ... and a shared-memory tree adds the 256 partial sums -- a fixed order again, so still deterministic.
template<class T>
__global__ void __launch_bounds__(256)
ttv_reduce_wide_kernel(const T* __restrict__ ws, T* __restrict__ c, uint64_t n, uint32_t ksplit, uint32_t accumulate, uint64_t stride)
{
  pdl_prologue();
  __shared__ T part[256];
  for (uint64_t j = blockIdx.x; j < n; j += gridDim.x) {
    T s = Num<T>::zero();
    for (uint32_t p = threadIdx.x; p < ksplit; p += 256) s = Num<T>::add(s, ws[(uint64_t)p * stride + j]);
    part[threadIdx.x] = s;
    __syncthreads();
    for (uint32_t h = 128; h > 0; h >>= 1) {
      if (threadIdx.x < h) part[threadIdx.x] = Num<T>::add(part[threadIdx.x], part[threadIdx.x + h]);
      __syncthreads();
    }
    if (threadIdx.x == 0) c[j] = accumulate ? Num<T>::add(c[j], part[0]) : part[0];
    __syncthreads();
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
  pdl_prologue();
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x)
    x[i] = synth<T>(seed, first + i);
}

} // namespace ttvb
