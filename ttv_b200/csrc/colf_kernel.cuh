// colf_kernel.cuh -- COLF: column GEMV whose rows are NARROWER than / not a multiple of a 16-byte vector, read as a flat
// stream of whole vectors, one WARP per slab or slab partition (sm_100a).
//
// A[outer][n_q][inner] with inner = 2, 3, 5, 6, 7, 9 ... elements: the asymmetric family of the reference (tiny leading
// extents; its slicing::small leaf is an n1 x n_q GEMV, tensor_times_vector.h:214, gemv_col matrix_times_vector.h:108-127).
// No 16-byte vector tiles such a row, so the column kernel loads 4 or 8 bytes per lane (2.3-4.3 TB/s for rows of 2 / 3 / 5 / 7
// floats).  But R = V / gcd(inner, V) consecutive rows -- a SUPER-ROW -- are L = inner / gcd(inner, V) whole vectors, and a
// slab is a contiguous run of super-rows.  So a warp streams its slab (or its partition of a long slab) flat: lane (ty, j)
// loads vector j of super-rows ty, ty + TY, ... -- consecutive lanes read consecutive vectors, up to 512 contiguous bytes per
// instruction, KU loads in flight per lane.  The V elements of that vector lie in at most two consecutive rows of the
// super-row, the same two for every super-row the lane visits: the first `sp` elements in row r0, the others in row r0 + 1.
// So the lane keeps V accumulators, one per element, and multiplies by b[R sr + r0] or b[R sr + r0 + 1]: two scalar loads
// that the lanes of a super-row share (L1 hits).  Afterwards the lanes of one phase j are added by a shuffle tree, lanes
// j < L hold a partial matrix [R][inner] (flat index V j + e), which goes through a warp-private strip of shared memory to
// be folded over its R rows.  No CTA-wide synchronisation anywhere.  A contraction cut across warps (ksplit) goes to the
// workspace [ksplit][outer * inner] and is finished by ttv_reduce_kernel / ttv_reduce_wide_kernel in fixed order.
// SHORT slabs (a few hundred bytes to a few KB: n = (2, 128, 2^21), (5, 64, 2^20)) would leave a lane with one or two
// loads and the warp with a shuffle tree per slab; there a warp works on SW slabs side by side, G = TY L lanes each, TY
// chosen so that a lane still has a batch of loads: the slabs are contiguous, so the warp keeps reading one contiguous
// run, the tree shrinks to log2(TY) levels, and the SW x inner outputs of the item are one contiguous run of C.
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
  uint32_t R, L, TY;        // rows / vectors per super-row, super-rows per step of a lane group (G = TY * L lanes per slab)
  uint32_t SW;              // slabs a warp works on side by side: SW * G <= 32 (SW > 1 only with ksplit == 1)
  uint32_t accumulate;      // only honoured when ksplit == 1
};

// STEP: the type the distances between the loads of a batch are computed in.  32 bits halve the address arithmetic, which
// is what a short slab spends its time on (8 388 608 slabs of 128 x 2 floats: 5 708 -> 5 760 GB/s, 262 144 of 256 x 3: 4 069 ->
// 4 355); long partitions measured 1.5-2.6 % FASTER with 64 bits (the loads of a batch leave in a different order), so each
// keeps its own -- 32 bits for slabs of at most two batches of 4-byte elements; 8-byte elements lose with 32 bits on short
// slabs too (4 194 304 slabs of 16 x 3 doubles: 5 294 -> 4 800) and always take 64 (tools/probe/tiny_inner.py, sessions 20-21).
template<class T, int KU, bool PRED, bool NA, class STEP>
__device__ __forceinline__ void colf_batch(T (&acc)[16 / sizeof(T)], const T* ap, const T* blo, const T* bhi, STEP astep, STEP bstep,
                                           uint64_t sr, STEP step, uint64_t n, uint32_t sp)
{
  constexpr int V = 16 / (int)sizeof(T);
  Vec<T, V> x[KU];
  T lo[KU], hi[KU];
#pragma unroll
  for (int s = 0; s < KU; ++s) {
    if constexpr (PRED) {
      const bool ok = sr + s * step < n;
      ld16_if<NA>(&x[s], ap + s * astep, ok);
      lo[s] = ok ? blo[s * bstep] : Num<T>::zero();
      hi[s] = ok ? bhi[s * bstep] : Num<T>::zero();
    } else {
      x[s]  = load_a<T, V>(ap + s * astep, NA);
      lo[s] = blo[s * bstep];
      hi[s] = bhi[s * bstep];
    }
  }
#pragma unroll
  for (int s = 0; s < KU; ++s)
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = Num<T>::madd(x[s].e[e], (uint32_t)e < sp ? lo[s] : hi[s], acc[e]);
}

// (8-byte elements: the KU vectors of A and the 2 KU elements of b of a batch are 64 registers; two CTAs per SM, no spills)
// NA: L1::no_allocate loads -- a lane group reads at least a 128-byte line per step; narrower groups reuse the line from L1.
template<class T, int KU, bool NA>
__global__ void __launch_bounds__(256, sizeof(T) == 8 ? 2 : 3)
ttv_colf_kernel(const ColfParams P)
{
  pdl_prologue();
  constexpr int V = 16 / (int)sizeof(T);
  __shared__ __align__(16) T strips[8][32 * V];                                   // per warp: SW partial matrices [R][inner] of L * V cells each

  const T* __restrict__ A = static_cast<const T*>(P.a);
  const T* __restrict__ B = static_cast<const T*>(P.b);
  T* __restrict__       C = static_cast<T*>(P.c);

  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t inner = (uint32_t)P.inner;
  const uint32_t G  = P.TY * P.L;                                   // lanes of one slab
  const uint32_t g  = lane / G, t = lane % G;                       // slab of the item, lane inside its group
  const uint32_t j  = t % P.L;
  const uint32_t ty = t / P.L;
  const uint32_t r0  = (V * j) / inner;                             // row of the vector's first element inside the super-row
  const uint32_t sp  = min((uint32_t)V, (r0 + 1) * inner - V * j);  // elements of the vector that lie in row r0
  const uint32_t rhi = min(r0 + 1, P.R - 1);
  T* strip = strips[warp];

  const uint64_t nsr   = P.nq / P.R;                                // whole super-rows of a slab
  const uint32_t astep = G * V, bstep = P.TY * P.R;                 // elements of A / of b between two steps of a lane (<= 128)
  const uint64_t slab_elems = P.nq * inner;
  const uint64_t ogroups = (P.outer + P.SW - 1) / P.SW;
  const uint64_t items = ogroups * P.ksplit;
  uint32_t span = 1;
  while (span < P.TY) span <<= 1;

  // what does not change from item to item: the cell of the strip a lane writes, the output a lane finishes first
  const bool holds_matrix = g < P.SW && t < P.L;                    // ty == 0: after the tree this lane holds V cells of [R][inner]
  Vec<T, V>* my_cells = reinterpret_cast<Vec<T, V>*>(strip) + (g * P.L + t);
  const uint32_t gg0 = lane / inner, c0 = lane % inner;             // output `lane` of an item: slab gg0, column c0
  const bool single = P.ksplit == 1;                                // no partition index to divide out (every short slab)
  const bool short32 = single && sizeof(T) == 4 && nsr <= 2ull * KU * P.TY;   // a slab is two batches at most: 32-bit steps

  for (uint64_t item = (uint64_t)blockIdx.x * 8 + warp; item < items; item += (uint64_t)gridDim.x * 8) {
    const uint64_t og = single ? item : item / P.ksplit;
    const uint32_t ks = single ? 0u : (uint32_t)(item - og * P.ksplit);
    const uint64_t o0 = og * P.SW;                                  // first slab of the item
    const uint64_t srbeg = single ? 0 : min((uint64_t)ks * P.srchunk, nsr);
    const uint64_t n = single ? nsr : min(srbeg + P.srchunk, nsr) - srbeg;
    const uint64_t o = o0 + g;

    T acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = Num<T>::zero();

    if (g < P.SW && o < P.outer) {
      const T* ap  = A + o * slab_elems + srbeg * (P.R * inner) + t * V;
      const T* blo = B + srbeg * P.R + (ty * P.R + r0);
      const T* bhi = B + srbeg * P.R + (ty * P.R + rhi);
      uint64_t sr = ty;
      if (short32) {
        for (; sr + (KU - 1) * P.TY < n; sr += KU * P.TY, ap += KU * astep, blo += KU * bstep, bhi += KU * bstep)
          colf_batch<T, KU, false, NA, uint32_t>(acc, ap, blo, bhi, astep, bstep, sr, P.TY, n, sp);
        if (sr < n) colf_batch<T, KU, true, NA, uint32_t>(acc, ap, blo, bhi, astep, bstep, sr, P.TY, n, sp);
      } else {
        const uint64_t astep64 = astep, bstep64 = bstep, ty64 = P.TY;
        for (; sr + (uint64_t)(KU - 1) * ty64 < n; sr += (uint64_t)KU * ty64, ap += KU * astep64, blo += KU * bstep64, bhi += KU * bstep64)
          colf_batch<T, KU, false, NA, uint64_t>(acc, ap, blo, bhi, astep64, bstep64, sr, ty64, n, sp);
        if (sr < n) colf_batch<T, KU, true, NA, uint64_t>(acc, ap, blo, bhi, astep64, bstep64, sr, ty64, n, sp);
      }
    }

    // lanes of one phase j of one slab: rows ty + h are folded onto ty (a source lane lies inside the same group)
    for (uint32_t h = span >> 1; h > 0; h >>= 1) {
      const bool take = ty < h && ty + h < P.TY;
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const T other = shfl_down_elem(acc[e], (int)(h * P.L));
        if (take) acc[e] = Num<T>::add(acc[e], other);
      }
    }
    if (holds_matrix) {
      Vec<T, V> cells;
#pragma unroll
      for (int e = 0; e < V; ++e) cells.e[e] = acc[e];
      *my_cells = cells;
    }
    __syncwarp();
    // the outputs of the item's slabs are one contiguous run of C
    const uint32_t outs = (uint32_t)min((uint64_t)P.SW, P.outer - o0) * inner;
    T* cout = C + ((single ? 0 : (uint64_t)ks * P.outer) + o0) * inner;
    for (uint32_t idx = lane; idx < outs; idx += 32) {
      const uint32_t gg = idx < 32 ? gg0 : idx / inner, c = idx < 32 ? c0 : idx % inner;
      const T* m = strip + gg * P.L * V;
      T val = m[c];
      for (uint32_t r = 1; r < P.R; ++r) val = Num<T>::add(val, m[c + r * inner]);
      if (ks + 1 == P.ksplit && nsr * P.R < P.nq) {                 // rows past the last whole super-row (only when outer == 1)
        for (uint64_t r = nsr * P.R; r < P.nq; ++r) val = Num<T>::madd(A[((o0 + gg) * P.nq + r) * inner + c], B[r], val);
      }
      cout[idx] = (P.accumulate && single) ? Num<T>::add(cout[idx], val) : val;
    }
    __syncwarp();                                                   // the strip is rewritten by the next item
  }
}

// ROWS OF TWO 4-byte elements (inner = 2: the leading extent 2 of the asymmetric family) are the one case where everything
// above is known at compile time: a vector is the two rows of its super-row (R = 2, L = 1), so the two elements of b are ONE
// aligned 8-byte load, no element needs a select, the two rows fold inside the lane, and after the shuffle tree lane 0 of a
// slab's group holds the slab's two outputs and stores them as 8 bytes -- no strip, no second pass over the lanes.  The
// general kernel spent 56 % of the issue slots on 8 388 608 slabs of 128 x 2 floats (ncu, 72 % gpu__dram_throughput).
template<class T, int KU, bool NA>
__global__ void __launch_bounds__(256, 3)
ttv_colf2_kernel(const ColfParams P)
{
  static_assert(sizeof(T) == 4, "rows of two 4-byte elements");
  pdl_prologue();
  using V4 = Vec<T, 4>;
  using V2 = Vec<T, 2>;
  const T* __restrict__ A = static_cast<const T*>(P.a);
  const T* __restrict__ B = static_cast<const T*>(P.b);
  T* __restrict__       C = static_cast<T*>(P.c);

  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t G = P.TY;                                          // lanes of one slab (L = 1)
  const uint32_t g = lane / G, ty = lane % G;
  const uint64_t nsr = P.nq / 2;
  const uint64_t slab_elems = P.nq * 2;
  const uint64_t ogroups = (P.outer + P.SW - 1) / P.SW;
  const uint64_t items = ogroups * P.ksplit;
  const bool single = P.ksplit == 1;
  uint32_t span = 1;
  while (span < P.TY) span <<= 1;

  for (uint64_t item = (uint64_t)blockIdx.x * 8 + warp; item < items; item += (uint64_t)gridDim.x * 8) {
    const uint64_t og = single ? item : item / P.ksplit;
    const uint32_t ks = single ? 0u : (uint32_t)(item - og * P.ksplit);
    const uint64_t srbeg = single ? 0 : min((uint64_t)ks * P.srchunk, nsr);
    const uint64_t n = single ? nsr : min(srbeg + P.srchunk, nsr) - srbeg;
    const uint64_t o = og * P.SW + g;
    const bool live = g < P.SW && o < P.outer;

    T acc[4] = {Num<T>::zero(), Num<T>::zero(), Num<T>::zero(), Num<T>::zero()};
    if (live) {
      const T* ap = A + o * slab_elems + (srbeg + ty) * 4;
      const T* bp = B + (srbeg + ty) * 2;
      const uint32_t astep = G * 4, bstep = G * 2;
      uint64_t sr = ty;
      for (; sr + (uint64_t)(KU - 1) * G < n; sr += (uint64_t)KU * G, ap += KU * astep, bp += KU * bstep) {
        V4 x[KU];
        V2 q[KU];
#pragma unroll
        for (int s = 0; s < KU; ++s) {
          x[s] = load_a<T, 4>(ap + s * astep, NA);
          q[s] = *reinterpret_cast<const V2*>(bp + s * bstep);
        }
#pragma unroll
        for (int s = 0; s < KU; ++s) {
          acc[0] = Num<T>::madd(x[s].e[0], q[s].e[0], acc[0]);
          acc[1] = Num<T>::madd(x[s].e[1], q[s].e[0], acc[1]);
          acc[2] = Num<T>::madd(x[s].e[2], q[s].e[1], acc[2]);
          acc[3] = Num<T>::madd(x[s].e[3], q[s].e[1], acc[3]);
        }
      }
      if (sr < n) {
        V4 x[KU];
        V2 q[KU];
#pragma unroll
        for (int s = 0; s < KU; ++s) {
          const bool ok = sr + (uint64_t)s * G < n;
          ld16_if<NA>(&x[s], ap + s * astep, ok);
          q[s] = ok ? *reinterpret_cast<const V2*>(bp + s * bstep) : zero_vec<T, 2>();
        }
#pragma unroll
        for (int s = 0; s < KU; ++s) {
          acc[0] = Num<T>::madd(x[s].e[0], q[s].e[0], acc[0]);
          acc[1] = Num<T>::madd(x[s].e[1], q[s].e[0], acc[1]);
          acc[2] = Num<T>::madd(x[s].e[2], q[s].e[1], acc[2]);
          acc[3] = Num<T>::madd(x[s].e[3], q[s].e[1], acc[3]);
        }
      }
    }
    T c0 = Num<T>::add(acc[0], acc[2]), c1 = Num<T>::add(acc[1], acc[3]);      // the two rows of the super-rows
    for (uint32_t h = span >> 1; h > 0; h >>= 1) {
      const T o0 = shfl_down_elem(c0, (int)h), o1 = shfl_down_elem(c1, (int)h);
      if (ty < h && ty + h < P.TY) { c0 = Num<T>::add(c0, o0); c1 = Num<T>::add(c1, o1); }
    }
    if (live && ty == 0) {
      if (ks + 1 == P.ksplit && (P.nq & 1)) {                       // an odd last row (only when outer == 1)
        const uint64_t r = P.nq - 1;
        c0 = Num<T>::madd(A[o * slab_elems + r * 2], B[r], c0);
        c1 = Num<T>::madd(A[o * slab_elems + r * 2 + 1], B[r], c1);
      }
      V2* out = reinterpret_cast<V2*>(C + ((single ? 0 : (uint64_t)ks * P.outer) + o) * 2);
      V2 val;
      if (P.accumulate && single) { const V2 old = *out; val.e[0] = Num<T>::add(old.e[0], c0); val.e[1] = Num<T>::add(old.e[1], c1); }
      else { val.e[0] = c0; val.e[1] = c1; }
      *out = val;
    }
  }
}

// SHORT slabs -- a slab is at most ONE batch of its lane group (n_q / R <= KU TY; always unsplit) -- leave the kernels above
// issue-bound: ncu counted 440 warp instructions per 4 KB item on 8 388 608 slabs of 128 x 2 floats, 127 of them IMAD (the
// addresses of A, b and C from scratch for every item), 43 ISETP, 27 BRA, against 24 loads and 64 FFMA / FSEL.  But nothing
// except the address of A changes from slab to slab: which of a lane's KU steps exist, and the elements of b that go with
// them.  This form therefore loads b ONCE per lane into registers, keeps the validity of the steps as per-lane constants
// (whole batches take plain loads), and walks A and C with one pointer increment per item.  PAIR is the rows-of-two form
// (see ttv_colf2_kernel): no selects, no strip.
template<class T, int KU, bool NA, bool PAIR>
__global__ void __launch_bounds__(256, sizeof(T) == 8 ? 2 : 3)
ttv_colfs_kernel(const ColfParams P)
{
  static_assert(!PAIR || sizeof(T) == 4, "rows of two 4-byte elements");
  pdl_prologue();
  constexpr int V = 16 / (int)sizeof(T);
  __shared__ __align__(16) T strips[PAIR ? 1 : 8][PAIR ? 4 : 32 * V];

  const T* __restrict__ A = static_cast<const T*>(P.a);
  const T* __restrict__ B = static_cast<const T*>(P.b);
  T* __restrict__       C = static_cast<T*>(P.c);

  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t inner = (uint32_t)P.inner;
  const uint32_t G  = P.TY * P.L;
  const uint32_t g  = lane / G, t = lane % G;
  const uint32_t j  = t % P.L, ty = t / P.L;
  const uint32_t r0  = (V * j) / inner;
  const uint32_t sp  = min((uint32_t)V, (r0 + 1) * inner - V * j);
  const uint32_t rhi = min(r0 + 1, P.R - 1);
  const uint32_t nsr = (uint32_t)(P.nq / P.R);                      // <= KU * TY
  const uint32_t astep = G * V;
  const uint64_t slab_elems = P.nq * inner;
  const bool in_group = g < P.SW;

  // the lane's steps and its elements of b: the same for every slab
  uint32_t okm = 0;                                                 // bit s: step s of this lane exists
  T lo[KU], hi[KU];
#pragma unroll
  for (int s = 0; s < KU; ++s) {
    const bool ok = in_group && ty + s * P.TY < nsr;
    okm |= ok ? 1u << s : 0u;
    lo[s] = ok ? B[(ty + s * P.TY) * P.R + r0] : Num<T>::zero();
    hi[s] = ok ? B[(ty + s * P.TY) * P.R + rhi] : Num<T>::zero();
  }
  const bool whole = (okm >> (KU - 1)) & 1u;                        // every step of this lane exists
  uint32_t span = 1;
  while (span < P.TY) span <<= 1;

  const uint64_t items = (P.outer + P.SW - 1) / P.SW;
  uint64_t item = (uint64_t)blockIdx.x * 8 + warp;
  const uint32_t istride = gridDim.x * 8;
  const T* ap = A + (item * P.SW + g) * slab_elems + t * V;         // (never dereferenced by a lane without a slab)
  const uint32_t astride = istride * P.SW * (uint32_t)slab_elems;   // a warp's item is at most 32 KU vectors: fits 32 bits
  const uint32_t cell = PAIR ? 0u : (g * P.L + t) * V;              // where the lane's V cells go in the warp's strip
  const uint32_t wrow = PAIR ? 0u : warp;

  for (; item < items; item += istride, ap += astride) {
    const uint64_t left = P.outer - item * P.SW;                    // slabs from the item's first one to the end
    const bool live = in_group && g < left;
    Vec<T, V> x[KU];
    if (whole && live) {
#pragma unroll
      for (int s = 0; s < KU; ++s) x[s] = load_a<T, V>(ap + s * astep, NA);
    } else {
#pragma unroll
      for (int s = 0; s < KU; ++s) ld16_if<NA>(&x[s], ap + s * astep, live && ((okm >> s) & 1u));
    }
    T acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = Num<T>::zero();
#pragma unroll
    for (int s = 0; s < KU; ++s)
#pragma unroll
      for (int e = 0; e < V; ++e) {
        if constexpr (PAIR) acc[e] = Num<T>::madd(x[s].e[e], e < 2 ? lo[s] : hi[s], acc[e]);
        else                acc[e] = Num<T>::madd(x[s].e[e], (uint32_t)e < sp ? lo[s] : hi[s], acc[e]);
      }
    T* cout = C + item * P.SW * inner;

    if constexpr (PAIR) {
      T c0 = Num<T>::add(acc[0], acc[2]), c1 = Num<T>::add(acc[1], acc[3]);
      for (uint32_t h = span >> 1; h > 0; h >>= 1) {
        const T o0 = shfl_down_elem(c0, (int)h), o1 = shfl_down_elem(c1, (int)h);
        if (ty < h && ty + h < P.TY) { c0 = Num<T>::add(c0, o0); c1 = Num<T>::add(c1, o1); }
      }
      if (live && ty == 0) {
        if (P.nq & 1) {                                             // an odd last row (only when outer == 1)
          const uint64_t r = P.nq - 1;
          c0 = Num<T>::madd(ap[r * 2], B[r], c0);
          c1 = Num<T>::madd(ap[r * 2 + 1], B[r], c1);
        }
        Vec<T, 2>* out = reinterpret_cast<Vec<T, 2>*>(cout + g * 2);
        Vec<T, 2> val;
        if (P.accumulate) { const Vec<T, 2> old = *out; val.e[0] = Num<T>::add(old.e[0], c0); val.e[1] = Num<T>::add(old.e[1], c1); }
        else { val.e[0] = c0; val.e[1] = c1; }
        *out = val;
      }
    } else {
      for (uint32_t h = span >> 1; h > 0; h >>= 1) {
        const bool take = ty < h && ty + h < P.TY;
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const T other = shfl_down_elem(acc[e], (int)(h * P.L));
          if (take) acc[e] = Num<T>::add(acc[e], other);
        }
      }
      if (in_group && t < P.L) {
        Vec<T, V> cells;
#pragma unroll
        for (int e = 0; e < V; ++e) cells.e[e] = acc[e];
        *reinterpret_cast<Vec<T, V>*>(&strips[wrow][cell]) = cells;
      }
      __syncwarp();
      const uint32_t outs = (uint32_t)min((uint64_t)P.SW, left) * inner;
      for (uint32_t idx = lane; idx < outs; idx += 32) {
        const uint32_t gg = idx / inner, c = idx - gg * inner;
        const uint32_t m = gg * P.L * V + c;
        T val = strips[wrow][m];
        for (uint32_t r = 1; r < P.R; ++r) val = Num<T>::add(val, strips[wrow][m + r * inner]);
        if ((uint64_t)nsr * P.R < P.nq) {                           // rows past the last whole super-row (only when outer == 1)
          for (uint64_t r = (uint64_t)nsr * P.R; r < P.nq; ++r) val = Num<T>::madd(A[((item * P.SW + gg) * P.nq + r) * inner + c], B[r], val);
        }
        cout[idx] = P.accumulate ? Num<T>::add(cout[idx], val) : val;
      }
      __syncwarp();
    }
  }
}

// TINY slabs of two-element rows: n_q = 2, 4, 8, 16, 32 rows of two 4-byte elements, i.e. a slab of G = n_q / 2 = 1 .. 16 vectors
// (16 .. 256 bytes; n = (2, 8, 2^25): the leading extent 2 with a small second mode).  With a lane per slab (what the short-slab
// form does from 128 bytes on) a load instruction touches 32 different lines and the L1 tag stage becomes the limit
// ([16777216, 16, 2]: 4 729 GB/s; the column kernel 1 814).  Here the lanes of a warp read CONSECUTIVE vectors -- G lanes per
// slab, 32 / G slabs per instruction, KU = 8 instructions = 4 KB in flight -- so a lane ends up with the partial sums of KU
// different slabs, one vector each, and the G partials of a slab sit in G neighbouring lanes.  They are added by a TRANSPOSING
// butterfly: at level m a lane sends the half of its values that its partner (lane ^ m) keeps and adds the half it receives
// to the half it keeps -- 4 + 2 + 1 shuffles per component instead of 3 x 8 -- and after log2(G) levels every lane holds the
// complete sums of 8 / G slabs, which it stores as 8 bytes each.  b (two elements per lane) lives in registers.
struct ColfTinyParams {
  const void* a;
  const void* b;
  void*       c;
  uint64_t outer;           // slabs
  uint32_t G;               // vectors (= lanes) per slab: 1, 2, 4, 8 or 16
  uint32_t accumulate;
};

template<class T>
__global__ void __launch_bounds__(256, 4)
ttv_colf_tiny_kernel(const ColfTinyParams P)
{
  static_assert(sizeof(T) == 4, "rows of two 4-byte elements");
  pdl_prologue();
  constexpr int KU = 8;
  using V4 = Vec<T, 4>;
  using V2 = Vec<T, 2>;
  const V4* __restrict__ A = static_cast<const V4*>(P.a);
  const T* __restrict__  B = static_cast<const T*>(P.b);
  T* __restrict__        C = static_cast<T*>(P.c);

  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t G = P.G, t = lane % G, g = lane / G, spw = 32u / G;     // vector inside the slab, slab inside a step, slabs per step
  const T b0 = B[2 * t], b1 = B[2 * t + 1];                              // the vector holds rows 2t and 2t + 1
  // the step whose complete sum ends up in slot i of this lane: i + the halves kept on the way (bit k of t keeps the upper half at level k)
  uint32_t sbase = 0;
  if (G > 1) sbase += (t & 1u) ? 4u : 0u;
  if (G > 2) sbase += (t & 2u) ? 2u : 0u;
  if (G > 4) sbase += (t & 4u) ? 1u : 0u;
  const uint32_t cnt = G >= 8 ? 1u : 8u / G;                             // complete sums per lane at the end
  const bool stores = G < 16 || !(t & 8u);

  const uint64_t per_item = (uint64_t)KU * spw;                          // slabs of one item (256 vectors, 4 KB)
  const uint64_t items = (P.outer + per_item - 1) / per_item;
  for (uint64_t item = (uint64_t)blockIdx.x * 8 + warp; item < items; item += (uint64_t)gridDim.x * 8) {
    const uint64_t slab0 = item * per_item;
    const V4* ap = A + item * (uint64_t)(KU * 32) + lane;
    V4 x[KU];
    if (slab0 + per_item <= P.outer) {
#pragma unroll
      for (int s = 0; s < KU; ++s) x[s] = load_a<T, 4>(reinterpret_cast<const T*>(ap + s * 32), true);
    } else {
#pragma unroll
      for (int s = 0; s < KU; ++s) ld16_if<true>(&x[s], ap + s * 32, slab0 + (uint64_t)s * spw + g < P.outer);
    }
    T c0[KU], c1[KU];
#pragma unroll
    for (int s = 0; s < KU; ++s) {
      c0[s] = Num<T>::madd(x[s].e[2], b1, Num<T>::madd(x[s].e[0], b0, Num<T>::zero()));
      c1[s] = Num<T>::madd(x[s].e[3], b1, Num<T>::madd(x[s].e[1], b0, Num<T>::zero()));
    }
    // transposing butterfly over the G lanes of a slab: 8 -> 4 -> 2 -> 1 values per lane
#pragma unroll
    for (int level = 0; level < 3; ++level) {
      const uint32_t m = 1u << level;
      constexpr int kHalf[3] = {4, 2, 1};
      const int half = kHalf[level];
      if (G > m) {
        const bool up = lane & m;
#pragma unroll
        for (int i = 0; i < half; ++i) {
          const T s0 = up ? c0[i] : c0[i + half], s1 = up ? c1[i] : c1[i + half];
          const T r0 = shfl_xor_elem(s0, (int)m), r1 = shfl_xor_elem(s1, (int)m);
          c0[i] = Num<T>::add(up ? c0[i + half] : c0[i], r0);
          c1[i] = Num<T>::add(up ? c1[i + half] : c1[i], r1);
        }
      }
    }
    if (G == 16) {                                                       // the two halves of a 16-lane slab
      c0[0] = Num<T>::add(c0[0], shfl_xor_elem(c0[0], 8));
      c1[0] = Num<T>::add(c1[0], shfl_xor_elem(c1[0], 8));
    }
    if (stores) {
#pragma unroll
      for (int i = 0; i < KU; ++i) {
        if ((uint32_t)i < cnt) {
          const uint64_t slab = slab0 + (uint64_t)(sbase + i) * spw + g;
          if (slab < P.outer) {
            V2* out = reinterpret_cast<V2*>(C + slab * 2);
            V2 val;
            if (P.accumulate) { const V2 old = *out; val.e[0] = Num<T>::add(old.e[0], c0[i]); val.e[1] = Num<T>::add(old.e[1], c1[i]); }
            else { val.e[0] = c0[i]; val.e[1] = c1[i]; }
            *out = val;
          }
        }
      }
    }
  }
}

} // namespace ttvb
