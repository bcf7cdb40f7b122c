// colr_kernel.cuh -- COLR: warp-autonomous column GEMV for rows off 16-byte boundaries, REALIGNED at load time.
//
// Same situation as colx_kernel.cuh (view A[outer][nq][inner] with inner % V != 0, short n_q): row (o, k) starts at
// phase phi = flat element mod V inside a 16-byte line.  A warp loads the ALIGNED vectors that cover its piece of a row
// (512 contiguous bytes per instruction); output column i of that row then sits at aligned position i + phi, i.e. in
// element (e + phi) % V of this lane's vector or of the right neighbour's.  COLW keeps one accumulator set per phase
// class (V*V registers per unit -> 115 registers, 2 CTAs per SM, 64 KB of loads in flight per SM; ncu: 24 % of the
// warps resident, long-scoreboard bound, 6.36 TB/s).  Here every loaded vector is shifted into place right away:
//
//     x[e] = (e + phi < V) ? mine[e + phi] : __shfl_down(mine[e + phi - V], 1)
//
// so a unit needs V accumulators, the kernel fits 80 registers and runs 3 CTAs per SM like the other kernels.  The
// shift costs phi shuffles per vector and NO selects, because the phases are compile-time constants: a batch covers KU
// consecutive rows with KU % V == 0, so row s of every batch of a tile has phase (f0 + s*inner) % V -- the tile's first
// row f0 picks one of V unrolled loop bodies (warp-uniform switch), and inner % V is a template parameter.
//
// Lane 31 only supplies the overlap vector, so a warp owns 31*V columns per unit.  b is read through L1 (one broadcast
// load per row).  The outputs of a warp leave through a shared-memory strip as whole 32-byte sectors (see the epilogue).  Replaces the same reference code as ttv_col_kernel (detail/matrix_times_vector.h:108-179 inside the
// loop nest of detail/tensor_times_vector.h:189-324).  Requires A 16-byte aligned; C needs only element alignment.
#pragma once

#include "kernels.cuh"

#include <utility>

namespace ttvb {

template<int... S, class F>
__device__ __forceinline__ void colr_static_for(std::integer_sequence<int, S...>, F&& f)
{
  (f(std::integral_constant<int, S>{}), ...);
}

// One batch: KU consecutive rows k .. k+KU-1 of NU units.  All lanes of the warp execute it together (shuffles).
// p0 = A + (flat element of row k) + (this lane's offset inside an aligned row, unit 0); bk = B + k.
template<class T, int V, int IM, int F0M, int NU, int KU, bool PRED>
__device__ __forceinline__ void colr_batch(T (&acc)[NU][V], const T* p0, const T* bk, uint64_t inner, uint32_t krem, int nvalid,
                                           const T* aend, bool stream)
{
  constexpr uint32_t WCOLS = 31 * V;
  Vec<T, V> v[NU][KU];
  T bb[KU];
  colr_static_for(std::make_integer_sequence<int, KU>{}, [&](auto sc) {
    constexpr int s  = decltype(sc)::value;
    constexpr int PH = (F0M + s * IM) % V;            // phase of row s of every batch of this tile
    const T* row = p0 + (uint64_t)s * inner - PH;     // this lane's aligned vector of row s, unit 0
    if constexpr (PRED) {
      const bool on = (uint32_t)s < krem;
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const T* p = row + u * WCOLS;
        if (on && u < nvalid) {
          if (p + V <= aend) v[u][s] = load_a<T, V>(p, stream);
          else {
#pragma unroll
            for (int e = 0; e < V; ++e) v[u][s].e[e] = (p + e < aend) ? p[e] : Num<T>::zero();
          }
        } else v[u][s] = zero_vec<T, V>();
      }
      bb[s] = on ? bk[s] : Num<T>::zero();
    } else {
#pragma unroll
      for (int u = 0; u < NU; ++u) v[u][s] = load_a<T, V>(row + u * WCOLS, stream);
      bb[s] = bk[s];
    }
  });
  colr_static_for(std::make_integer_sequence<int, KU>{}, [&](auto sc) {
    constexpr int s  = decltype(sc)::value;
    constexpr int PH = (F0M + s * IM) % V;
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      T nx[V];                                         // the first PH elements of the right neighbour's vector
#pragma unroll
      for (int e = 0; e < PH; ++e) nx[e] = shfl_down_elem(v[u][s].e[e], 1);
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const T x = (e + PH < V) ? v[u][s].e[(e + PH) % V] : nx[(e + PH) % V];
        acc[u][e] = Num<T>::madd(x, bb[s], acc[u][e]);
      }
    }
  });
}

template<class T, int V, int IM, int F0M, int NU, int KU>
__device__ __forceinline__ void colr_rows(T (&acc)[NU][V], const T* A, const T* B, uint64_t f0, uint64_t inner, uint64_t voff,
                                          uint64_t kbeg, uint64_t kend, int nvalid, const T* aend, bool stream)
{
  constexpr uint32_t WCOLS = 31 * V;
  const T* p0 = A + f0 + voff;
  const T* bk = B + kbeg;
  uint32_t krem = (uint32_t)(kend - kbeg);
  // full batches: every lane has all its units and the last load of the batch (row k+KU-1, unit NU-1) is inside A
  const bool whole = __all_sync(0xffffffffu, nvalid == NU);
  if (whole)
    for (; krem >= (uint32_t)KU; krem -= KU, p0 += (uint64_t)KU * inner, bk += KU) {
      const T* last = p0 + (uint64_t)(KU - 1) * inner + (uint64_t)(NU - 1) * WCOLS;      // at most V-1 before the load
      if (!__all_sync(0xffffffffu, last + V <= aend)) break;
      colr_batch<T, V, IM, F0M, NU, KU, false>(acc, p0, bk, inner, krem, nvalid, aend, stream);
    }
  for (; krem > 0; krem = krem > (uint32_t)KU ? krem - KU : 0, p0 += (uint64_t)KU * inner, bk += KU)
    colr_batch<T, V, IM, F0M, NU, KU, true>(acc, p0, bk, inner, krem, nvalid, aend, stream);
}

template<class T, int V, int IM, int NU, int KU>
__global__ void __launch_bounds__(256, 3)
ttv_colr_kernel(const TileParams P)
{
  pdl_prologue();
  static_assert(V > 1 && V <= 4 && (V & (V - 1)) == 0 && KU % V == 0, "COLR walks whole phase periods");
  const T* __restrict__ A = static_cast<const T*>(P.a);
  const T* __restrict__ B = static_cast<const T*>(P.b);
  T* __restrict__       C = static_cast<T*>(P.c);
  const T* aend = A + P.outer * P.nq * P.inner;
  extern __shared__ __align__(16) unsigned char smem_raw[];    // [warps][NU * 31 * V]: output strips

  const uint32_t lane = threadIdx.x & 31u;
  const bool     stream = P.stream != 0;
  constexpr uint32_t WCOLS = 31 * V;                           // columns a warp owns per unit
  const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);

  for (uint64_t tile = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < P.tiles; tile += warps) {
    const uint64_t it = tile % P.itiles;
    const uint64_t r  = tile / P.itiles;
    const uint32_t ks = (uint32_t)(r % P.ksplit);
    const uint64_t o  = r / P.ksplit;
    const uint64_t kbeg = (uint64_t)ks * P.kchunk;
    const uint64_t kend = min(kbeg + P.kchunk, P.nq);
    const uint64_t c0 = it * (uint64_t)(NU * WCOLS);
    const uint64_t f0 = (o * P.nq + kbeg) * P.inner;           // flat element of row kbeg

    // unit u owns columns [cu, min(cu + WCOLS, inner)); this lane loads vector cu/V + lane of every row
    int nvalid = 0;
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const uint64_t cu = c0 + (uint64_t)u * WCOLS;
      if (cu < P.inner && cu / V + lane <= (min(cu + WCOLS, P.inner) + V - 2) / V) nvalid = u + 1;
    }

    T acc[NU][V];
#pragma unroll
    for (int u = 0; u < NU; ++u)
#pragma unroll
      for (int e = 0; e < V; ++e) acc[u][e] = Num<T>::zero();

    const uint64_t voff = c0 + (uint64_t)lane * V;             // this lane's offset inside an aligned row, unit 0
    switch ((uint32_t)(f0 & (V - 1))) {                        // warp-uniform: the tile's first row fixes all phases
      case 0: colr_rows<T, V, IM, 0, NU, KU>(acc, A, B, f0, P.inner, voff, kbeg, kend, nvalid, aend, stream); break;
      case 1: colr_rows<T, V, IM, 1, NU, KU>(acc, A, B, f0, P.inner, voff, kbeg, kend, nvalid, aend, stream); break;
      default:
        if constexpr (V > 2) {
          if ((f0 & (V - 1)) == 2) colr_rows<T, V, IM, 2, NU, KU>(acc, A, B, f0, P.inner, voff, kbeg, kend, nvalid, aend, stream);
          else                     colr_rows<T, V, IM, 3, NU, KU>(acc, A, B, f0, P.inner, voff, kbeg, kend, nvalid, aend, stream);
        }
        break;
    }

    // Epilogue.  A lane's V outputs start at an arbitrary element of C, so storing them directly would write partial
    // 32-byte sectors -- and L2 answers every partial-sector write with a DRAM read of that sector (ncu: dram reads =
    // |A| + |C| instead of |A|).  The warp's outputs are one contiguous run of C, so they go through a warp-private
    // strip of shared memory and leave lane-contiguously from a sector boundary: every store instruction covers whole
    // sectors except at the two ends of the run.
    T* strip = reinterpret_cast<T*>(smem_raw) + (size_t)(threadIdx.x >> 5) * (NU * WCOLS);
    if (lane < 31) {
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        Vec<T, V> val;
#pragma unroll
        for (int e = 0; e < V; ++e) val.e[e] = acc[u][e];
        *reinterpret_cast<Vec<T, V>*>(strip + u * WCOLS + lane * V) = val;
      }
    }
    __syncwarp();
    if (c0 < P.inner) {
      T* dst = C + (P.ksplit > 1 ? (uint64_t)ks * P.outer * P.inner : 0) + o * P.inner + c0;
      const int64_t n = (int64_t)(min(c0 + (uint64_t)(NU * WCOLS), P.inner) - c0);
      constexpr uint32_t SECT = 32 / sizeof(T) > 0 ? 32 / sizeof(T) : 1;             // elements per sector
      const int64_t mis = (int64_t)((reinterpret_cast<uintptr_t>(dst) / sizeof(T)) % SECT);
      for (int64_t w = (int64_t)lane - mis; w < n; w += 32)
        if (w >= 0) dst[w] = (P.accumulate && P.ksplit == 1) ? Num<T>::add(dst[w], strip[w]) : strip[w];
    }
    __syncwarp();
  }
}

} // namespace ttvb
