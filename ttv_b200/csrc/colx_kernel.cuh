// colx_kernel.cuh -- COLX: column GEMV for rows that do NOT start on 16-byte boundaries (odd inner extents).
//
// View A[outer][nq][inner] with inner % V != 0 (1625 floats, 23^3 floats, 21^6 doubles ...): row (o, k) starts at flat
// element (o*nq + k)*inner, i.e. at phase phi = that mod V inside a 16-byte line, and the phase changes from row to row.
// The plain column kernel then has to fall back to 4-/8-byte loads.  Two facts rescue the 16-byte loads:
//   * rows are contiguous, so the aligned vectors that cover a row are all inside A (they merely contain up to V-1
//     elements of the neighbouring rows at either end);
//   * rows k, k+V, k+2V, ... of one slab share the same phase, because V*inner is a multiple of V.
// So the CTA puts TY = (a multiple of) V lanes along n_q: lane ty visits rows k = ty, ty+TY, ... which all have phase
// phi(ty), loads ALIGNED vectors j = c0/V + tx (+ u*TX) of each row, and accumulates them element-wise.  Element e of
// its vector j is output column j*V + e - phi(ty).  The main loop is the plain column kernel's (no shuffles, no
// predicates); the phase only enters the epilogue, where the TY partial sums are scattered to their columns in shared
// memory, summed in fixed order and stored.  A tile owns W = (TX*NU - 1)*V output columns and loads TX*NU vectors per
// row: the one extra vector is the cost of the shifted phases (it is shared with the neighbouring tile, < 2 % of the
// traffic and an L2 hit).
//
// Replaces the same reference code as ttv_col_kernel (detail/matrix_times_vector.h:108-179 inside the loop nest of
// detail/tensor_times_vector.h:189-324).  Requires A 16-byte aligned; C needs only element alignment.
#pragma once

#include "kernels.cuh"

namespace ttvb {

// guarded variant of col_batch for the edges: predicated on unit / k, and never reads at or beyond `aend`
template<class T, int V, int NU, int KU>
__device__ __forceinline__ void colx_batch_edge(T (&acc)[NU][V], const T* ap, uint64_t a_ustride, uint64_t kstride, const T* sb,
                                                uint32_t k, uint32_t tyn, uint32_t kn, int nvalid, const T* aend, bool stream)
{
  Vec<T, V> v[NU][KU];
#pragma unroll
  for (int u = 0; u < NU; ++u)
#pragma unroll
    for (int s = 0; s < KU; ++s) {
      const T* p = ap + u * a_ustride + s * kstride;
      if (u < nvalid && k + s * tyn < kn) {
        if (p + V <= aend) v[u][s] = load_a<T, V>(p, stream);
        else {
#pragma unroll
          for (int e = 0; e < V; ++e) v[u][s].e[e] = (p + e < aend) ? p[e] : Num<T>::zero();
        }
      } else v[u][s] = zero_vec<T, V>();
    }
#pragma unroll
  for (int s = 0; s < KU; ++s) {
    const T bb = (k + s * tyn < kn) ? sb[k + s * tyn] : Num<T>::zero();
#pragma unroll
    for (int u = 0; u < NU; ++u)
#pragma unroll
      for (int j = 0; j < V; ++j) acc[u][j] = Num<T>::madd(v[u][s].e[j], bb, acc[u][j]);
  }
}

template<class T, int V, int NU, int KU>
__global__ void __launch_bounds__(256, min_ctas<NU, KU>())
ttv_colx_kernel(const TileParams P)
{
  pdl_prologue();
  static_assert(V > 1 && (V & (V - 1)) == 0, "COLX is for vector loads");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sb  = reinterpret_cast<T*>(smem_raw);          // [kb]
  T* red = sb + P.kb;                               // [ty][wcols]

  const T* __restrict__ A = static_cast<const T*>(P.a);
  const T* __restrict__ B = static_cast<const T*>(P.b);
  T* __restrict__       C = static_cast<T*>(P.c);
  const T* aend = A + P.outer * P.nq * P.inner;

  const uint32_t tid = threadIdx.x;
  const uint32_t tx  = tid % P.tx;
  const uint32_t ty  = tid / P.tx;
  const bool     live = ty < P.ty;
  const bool     stream = P.stream != 0;
  const uint32_t W = (uint32_t)P.c_ustride;              // output columns of one tile
  const uint64_t kstride = (uint64_t)P.ty * P.inner;
  const bool     b_resident = (P.ksplit == 1) && (P.nq <= P.kb);
  // elements past `ap` the last load of a full batch reaches
  const uint64_t reach = (uint64_t)(NU - 1) * P.a_ustride + (uint64_t)(KU - 1) * kstride + V;

  if (b_resident) {
    for (uint32_t j = tid; j < (uint32_t)P.nq; j += blockDim.x) sb[j] = B[j];
    __syncthreads();
  }

  for (uint64_t tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
    const uint64_t it = tile % P.itiles;
    const uint64_t r  = tile / P.itiles;
    const uint32_t ks = (uint32_t)(r % P.ksplit);
    const uint64_t o  = r / P.ksplit;
    const uint64_t c0 = it * W;                               // columns [c0, c1) belong to this tile (c0 % V == 0)
    const uint64_t c1 = min(c0 + W, P.inner);
    const uint64_t kbeg = (uint64_t)ks * P.kchunk;
    const uint64_t kend = min(kbeg + P.kchunk, P.nq);

    // this lane's rows kbeg+ty, kbeg+ty+TY, ... all start at the same phase inside a 16-byte line
    const uint64_t f0  = (o * P.nq + kbeg + ty) * P.inner;
    const uint32_t phi = (uint32_t)(f0 & (V - 1));
    const uint64_t j0  = c0 / V + tx;                         // vector of unit 0 inside the (aligned) row
    // last vector any phase needs for this tile; lanes whose phase does not need it load it anyway (it is inside A
    // and the epilogue drops it), so that whole warps stay on the unpredicated path
    const uint64_t jlast = (c1 + V - 2) / V;
    int nvalid = 0;
    if (live && j0 <= jlast) {
      const uint64_t room = (jlast - j0) / P.tx + 1;
      nvalid = room < (uint64_t)NU ? (int)room : NU;
    }

    T acc[NU][V];
#pragma unroll
    for (int u = 0; u < NU; ++u)
#pragma unroll
      for (int j = 0; j < V; ++j) acc[u][j] = Num<T>::zero();

    for (uint64_t k0 = kbeg; k0 < kend; k0 += P.kb) {
      const uint32_t kn = (uint32_t)min((uint64_t)P.kb, kend - k0);
      if (!b_resident) {
        __syncthreads();
        for (uint32_t j = tid; j < kn; j += blockDim.x) sb[j] = B[k0 + j];
        __syncthreads();
      }
      if (nvalid > 0) {
        const T* ap = A + (f0 - phi) + (k0 - kbeg) * P.inner + j0 * V;
        uint32_t k = ty;
        if (nvalid == NU)
          for (; k + (KU - 1) * P.ty < kn && ap + reach <= aend; k += KU * P.ty, ap += KU * kstride)      // full batches
            col_batch<T, V, NU, KU, false>(acc, ap, P.a_ustride, kstride, sb, k, P.ty, kn, nvalid, stream);
        for (; k < kn; k += KU * P.ty, ap += KU * kstride)                                                // edges
          colx_batch_edge<T, V, NU, KU>(acc, ap, P.a_ustride, kstride, sb, k, P.ty, kn, nvalid, aend, stream);
      }
    }

    // epilogue: scatter the partial sums to their columns, add the TY phases in fixed order, store
    __syncthreads();
    if (live) {
#pragma unroll
      for (int u = 0; u < NU; ++u)
        if (u < nvalid) {
          const int64_t col = (int64_t)((j0 + (uint64_t)u * P.tx) * V) - (int64_t)phi - (int64_t)c0;
#pragma unroll
          for (int e = 0; e < V; ++e) {
            const int64_t i = col + e;
            if (i >= 0 && i < (int64_t)(c1 - c0)) red[(size_t)ty * W + i] = acc[u][e];
          }
        }
    }
    __syncthreads();
    T* dst = C + (P.ksplit > 1 ? (uint64_t)ks * P.outer * P.inner : 0) + o * P.inner + c0;
    for (uint32_t i = tid; i < (uint32_t)(c1 - c0); i += blockDim.x) {
      T s = red[i];
      for (uint32_t y = 1; y < P.ty; ++y) s = Num<T>::add(s, red[(size_t)y * W + i]);
      if (P.accumulate && P.ksplit == 1) s = Num<T>::add(dst[i], s);
      dst[i] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// COLW: the warp-autonomous form of the same idea.  One thread keeps all V phase classes to itself -- rows k with
// k % V == c accumulate into acc[c] -- so it walks CONSECUTIVE rows (KU of them in flight), needs no shared memory and
// no __syncthreads: a warp loads aligned vectors j = c0/V + lane of every row (512 contiguous bytes), and at the end
// each lane assembles its V output columns c0 + V*lane .. from its own and its right neighbour's accumulators with one
// shuffle per register (column i of a row with phase phi sits at aligned position i + phi).  Lane 31 only supplies
// the overlap vector, so a warp owns 31*V columns per unit.  b is read through L1 (one broadcast load per row).
// ------------------------------------------------------------------------------------------------------------------
#ifndef TTVB_COLW_CTAS
#define TTVB_COLW_CTAS 2
#endif

template<class T, int V>
__device__ __forceinline__ void colw_shift_add(T (&out)[V], const T (&mine)[V], const T (&next)[V], uint32_t phi)
{
  // out[e] += cat[e + phi], cat = mine ++ next; phi < V is warp-uniform per class, indices are static per case
#pragma unroll
  for (int ph = 0; ph < V; ++ph)
    if (phi == (uint32_t)ph) {
#pragma unroll
      for (int e = 0; e < V; ++e) out[e] = Num<T>::add(out[e], (e + ph < V) ? mine[(e + ph) % V] : next[(e + ph) % V]);
    }
}

template<class T, int V, int NU, int KU>
__global__ void __launch_bounds__(256, TTVB_COLW_CTAS)
ttv_colw_kernel(const TileParams P)
{
  pdl_prologue();
  static_assert(V > 1 && (V & (V - 1)) == 0 && KU % V == 0, "COLW walks whole phase periods");
  const T* __restrict__ A = static_cast<const T*>(P.a);
  const T* __restrict__ B = static_cast<const T*>(P.b);
  T* __restrict__       C = static_cast<T*>(P.c);
  const T* aend = A + P.outer * P.nq * P.inner;

  const uint32_t lane = threadIdx.x & 31u;
  const bool     stream = P.stream != 0;
  constexpr uint32_t WCOLS = 31 * V;                           // columns a warp owns per unit
  const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);

  for (uint64_t tile = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < P.tiles; tile += warps) {
    const uint64_t it = tile % P.itiles;
    const uint64_t r  = tile / P.itiles;
    const uint32_t ks = (uint32_t)(r % P.ksplit);
    const uint64_t o  = r / P.ksplit;
    const uint64_t kbeg = (uint64_t)ks * P.kchunk;             // multiple of V
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

    T acc[NU][V][V];
#pragma unroll
    for (int u = 0; u < NU; ++u)
#pragma unroll
      for (int c = 0; c < V; ++c)
#pragma unroll
        for (int e = 0; e < V; ++e) acc[u][c][e] = Num<T>::zero();

    const uint64_t voff = c0 + (uint64_t)lane * V;             // this lane's offset inside an aligned row, unit 0
    for (uint64_t k = kbeg; k < kend; k += KU) {
      Vec<T, V> v[NU][KU];
      T bb[KU];
      const uint64_t fk = f0 + (k - kbeg) * P.inner;
      // the last load of a full batch: row k+KU-1, unit NU-1
      const T* last = A + ((fk + (uint64_t)(KU - 1) * P.inner) & ~(uint64_t)(V - 1)) + voff + (uint64_t)(NU - 1) * WCOLS;
      if (nvalid == NU && k + KU <= kend && last + V <= aend) {
#pragma unroll
        for (int s = 0; s < KU; ++s) {
          const T* row = A + ((fk + (uint64_t)s * P.inner) & ~(uint64_t)(V - 1)) + voff;
#pragma unroll
          for (int u = 0; u < NU; ++u) v[u][s] = load_a<T, V>(row + u * WCOLS, stream);
          bb[s] = B[k + s];
        }
      } else {
#pragma unroll
        for (int s = 0; s < KU; ++s) {
          const T* row = A + ((fk + (uint64_t)s * P.inner) & ~(uint64_t)(V - 1)) + voff;
          const bool on = k + s < kend;
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
          bb[s] = on ? B[k + s] : Num<T>::zero();
        }
      }
#pragma unroll
      for (int s = 0; s < KU; ++s)
#pragma unroll
        for (int u = 0; u < NU; ++u)
#pragma unroll
          for (int e = 0; e < V; ++e) acc[u][s % V][e] = Num<T>::madd(v[u][s].e[e], bb[s], acc[u][s % V][e]);
    }

    // epilogue: classes in fixed order; class c has phase (f0 + c*inner) % V
    T* dst = C + (P.ksplit > 1 ? (uint64_t)ks * P.outer * P.inner : 0) + o * P.inner;
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      T out[V];
#pragma unroll
      for (int e = 0; e < V; ++e) out[e] = Num<T>::zero();
#pragma unroll
      for (int c = 0; c < V; ++c) {
        const uint32_t phi = (uint32_t)((f0 + (uint64_t)c * P.inner) & (V - 1));
        T next[V];
#pragma unroll
        for (int e = 0; e < V; ++e) next[e] = shfl_down_elem(acc[u][c][e], 1);
        colw_shift_add<T, V>(out, acc[u][c], next, phi);
      }
      const uint64_t cu = c0 + (uint64_t)u * WCOLS;
      const uint64_t cend = min(cu + WCOLS, P.inner);
      if (lane < 31) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const uint64_t col = cu + (uint64_t)lane * V + e;
          if (col < cend) dst[col] = (P.accumulate && P.ksplit == 1) ? Num<T>::add(dst[col], out[e]) : out[e];
        }
      }
    }
  }
}

} // namespace ttvb
