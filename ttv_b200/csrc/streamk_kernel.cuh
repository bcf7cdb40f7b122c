// streamk_kernel.cuh -- STREAMK: a tiny odd inner extent under a LONG contraction (sm_100a).
//
// A[outer][n_q][inner] with inner = 3, 5, 6, 7, 9 ... elements and n_q in the thousands to millions (the asymmetric family of
// the reference: tiny leading extents, one huge mode; its slicing::small leaf is an n1 x n_q GEMV, tensor_times_vector.h:214).
// A row is 12-48 bytes, so no 16-byte vector tiles it and the column kernel falls back to 4- or 8-byte loads with its lanes
// strung along n_q: 2.9-4.3 TB/s for 4-byte elements.  But a slab A[o] is ONE contiguous run of n_q * inner elements and b is
// contiguous too, so -- as in STREAM (stream_kernel.cuh) -- a CTA lets the TMA unit copy whole runs of rows into shared
// memory (cp.async.bulk, rounded outwards to 16 bytes) together with the matching piece of b, three stages deep, and
// computes from there.  Unlike STREAM the outputs of a slab are only `inner` values, so the THREADS run along n_q: thread t
// takes rows t, t + 256, ... of a stage and keeps `inner` accumulators; at the end of its k-range the CTA adds the 256 partial
// vectors in a shared-memory tree (fixed order: deterministic) and writes `inner` outputs -- to C, or to the split-n_q
// workspace [ksplit][outer * inner] when the contraction is cut across CTAs (ttv_reduce_kernel / ttv_reduce_wide_kernel add
// the partitions).
//
// Replaces, for this regime, gemv_col inside the loop nest (detail/matrix_times_vector.h:108-127,
// detail/tensor_times_vector.h:189-324).
#pragma once

#include "stream_kernel.cuh"

namespace ttvb {

struct StreamkParams {
  const void* a;
  const void* b;
  void*       c;              // C, or the workspace [ksplit][outer * inner] when ksplit > 1
  uint64_t outer, nq, inner;
  uint64_t kchunk;            // rows per partition (a multiple of rows_per_stage)
  uint64_t total_bytes_a, total_bytes_b;
  uint32_t ksplit;
  uint32_t rows_per_stage;
  uint32_t a_stage_bytes, b_stage_bytes;     // bytes of one stage for the rows / for the piece of b (multiples of 128)
  uint32_t accumulate;        // only honoured when ksplit == 1
};

template<class T, int NS, int IMAX>
__global__ void __launch_bounds__(256, 2)
ttv_streamk_kernel(const StreamkParams P)
{
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* a_st = smem_raw;                                                        // [NS][a_stage_bytes]
  unsigned char* b_st = smem_raw + (size_t)NS * P.a_stage_bytes;                         // [NS][b_stage_bytes]
  T*        red  = reinterpret_cast<T*>(b_st + (size_t)NS * P.b_stage_bytes);            // [256][inner]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(red) + (((size_t)256 * P.inner * sizeof(T) + 15) & ~(size_t)15));

  const unsigned char* __restrict__ Ab = static_cast<const unsigned char*>(P.a);
  const unsigned char* __restrict__ Bb = static_cast<const unsigned char*>(P.b);
  const T* __restrict__ A = static_cast<const T*>(P.a);
  const T* __restrict__ B = static_cast<const T*>(P.b);
  T* __restrict__       C = static_cast<T*>(P.c);

  const uint32_t tid = threadIdx.x;
  const uint32_t inner = (uint32_t)P.inner;
  const uint64_t a_tail = P.total_bytes_a & ~(uint64_t)15, b_tail = P.total_bytes_b & ~(uint64_t)15;   // what bulk copies may touch

  if (tid == 0) {
#pragma unroll
    for (int st = 0; st < NS; ++st) tma::mbar_init(&bars[st], 1);
    tma::fence_barrier_init();
  }
  __syncthreads();

  uint32_t it = 0;                                       // stages consumed so far by this CTA (ring position and parity)
  const uint64_t tiles = P.outer * P.ksplit;
  for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const uint64_t o  = tile / P.ksplit;
    const uint32_t ks = (uint32_t)(tile % P.ksplit);
    const uint64_t kbeg = (uint64_t)ks * P.kchunk, kend = min(kbeg + P.kchunk, P.nq);
    const uint32_t chunks = kbeg < kend ? (uint32_t)((kend - kbeg + P.rows_per_stage - 1) / P.rows_per_stage) : 0u;

    // rows [k0, k0 + n) of slab o and b[k0 .. k0 + n): two contiguous byte ranges, each rounded outwards to 16 bytes
    auto issue = [&](uint32_t ch, uint32_t st) {
      const uint64_t k0 = kbeg + (uint64_t)ch * P.rows_per_stage;
      const uint64_t n  = min((uint64_t)P.rows_per_stage, kend - k0);
      const uint64_t ag0 = ((o * P.nq + k0) * inner) * sizeof(T), ag1 = ag0 + n * inner * sizeof(T);
      const uint64_t bg0 = k0 * sizeof(T), bg1 = bg0 + n * sizeof(T);
      const uint64_t alo = ag0 & ~(uint64_t)15, blo = bg0 & ~(uint64_t)15;
      const uint64_t ahi = min((ag1 + 15) & ~(uint64_t)15, a_tail), bhi = min((bg1 + 15) & ~(uint64_t)15, b_tail);
      const uint32_t abytes = ahi > alo ? (uint32_t)(ahi - alo) : 0u, bbytes = bhi > blo ? (uint32_t)(bhi - blo) : 0u;
      if (abytes + bbytes) {
        tma::mbar_expect_tx(&bars[st], abytes + bbytes);
        if (abytes) tma::bulk_g2s(a_st + (size_t)st * P.a_stage_bytes, Ab + alo, abytes, &bars[st]);
        if (bbytes) tma::bulk_g2s(b_st + (size_t)st * P.b_stage_bytes, Bb + blo, bbytes, &bars[st]);
      } else {
        tma::mbar_arrive(&bars[st]);
      }
    };

    if (tid == 0) {
      tma::fence_proxy_async();                          // the stages were read with ordinary loads by the tile before
      for (uint32_t ch = 0; ch < chunks && ch < (uint32_t)NS; ++ch) issue(ch, (it + ch) % NS);
    }

    T acc[IMAX];
#pragma unroll
    for (int c = 0; c < IMAX; ++c) acc[c] = Num<T>::zero();

    for (uint32_t ch = 0; ch < chunks; ++ch, ++it) {
      const uint32_t st = it % NS, parity = (it / NS) & 1u;
      tma::mbar_wait(&bars[st], parity);
      const uint64_t k0 = kbeg + (uint64_t)ch * P.rows_per_stage;
      const uint32_t n  = (uint32_t)min((uint64_t)P.rows_per_stage, kend - k0);
      const uint64_t ag0 = ((o * P.nq + k0) * inner) * sizeof(T), ag1 = ag0 + (uint64_t)n * inner * sizeof(T);
      const uint64_t bg0 = k0 * sizeof(T), bg1 = bg0 + (uint64_t)n * sizeof(T);
      T* rows = reinterpret_cast<T*>(a_st + (size_t)st * P.a_stage_bytes + (ag0 - (ag0 & ~(uint64_t)15)));
      T* bs   = reinterpret_cast<T*>(b_st + (size_t)st * P.b_stage_bytes + (bg0 - (bg0 & ~(uint64_t)15)));
      // the last < 16 bytes of A and of b are not covered by bulk copies: plain loads
      if (ag1 > a_tail || bg1 > b_tail) {
        if (ag1 > a_tail) {
          const uint64_t from = max(a_tail, ag0);
          const uint32_t cnt = (uint32_t)((ag1 - from) / sizeof(T));
          if (tid < cnt) rows[(from - ag0) / sizeof(T) + tid] = A[from / sizeof(T) + tid];
        }
        if (bg1 > b_tail) {
          const uint64_t from = max(b_tail, bg0);
          const uint32_t cnt = (uint32_t)((bg1 - from) / sizeof(T));
          if (tid < cnt) bs[(from - bg0) / sizeof(T) + tid] = B[from / sizeof(T) + tid];
        }
        __syncthreads();
      }

      for (uint32_t r = tid; r < n; r += 256) {
        const T bb = bs[r];
        const T* row = rows + (size_t)r * inner;
#pragma unroll
        for (int c = 0; c < IMAX; ++c)
          if (c < (int)inner) acc[c] = Num<T>::madd(row[c], bb, acc[c]);
      }

      __syncthreads();                                   // everybody is done with this stage
      if (tid == 0 && ch + NS < chunks) {
        tma::fence_proxy_async();
        issue(ch + NS, st);
      }
    }

    // the 256 partial vectors of this (slab, partition): shared-memory tree in fixed order
    T* mine = red + (size_t)tid * inner;
#pragma unroll
    for (int c = 0; c < IMAX; ++c)
      if (c < (int)inner) mine[c] = acc[c];
    __syncthreads();
    for (uint32_t h = 128; h > 0; h >>= 1) {
      if (tid < h) {
        const T* other = mine + (size_t)h * inner;
        for (uint32_t c = 0; c < inner; ++c) mine[c] = Num<T>::add(mine[c], other[c]);
      }
      __syncthreads();
    }
    if (tid < inner) {
      T* out = C + ((P.ksplit > 1 ? (uint64_t)ks * P.outer : 0) + o) * inner + tid;
      *out = (P.accumulate && P.ksplit == 1) ? Num<T>::add(*out, red[tid]) : red[tid];
    }
    __syncthreads();                                     // red is reused by the next tile
  }
}

} // namespace ttvb
