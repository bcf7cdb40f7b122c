// stream_kernel.cuh -- STREAM kernel: small slabs staged through shared memory with TMA bulk copies (sm_100a).
//
// When a slab A[o][:][:] (n_q x inner elements) is small and of odd size -- 23 x 23 floats, fibers of 21 doubles --
// no thread mapping gives aligned 16-byte global loads: slab and row starts fall anywhere inside a line.  But slabs
// are contiguous, so a CTA can treat S consecutive slabs as ONE contiguous byte range, round it outwards to 16 bytes
// and let the TMA unit copy it into shared memory with a single cp.async.bulk per stage (SASS: UBLKCP), completion
// signalled on an mbarrier.  Global traffic is then perfectly coalesced regardless of n_q and inner, no registers or
// issue slots are spent on loads, and NS stages keep NS * stage_bytes in flight per CTA.  The arithmetic reads the
// slab from shared memory: thread u owns output (slab, i) and walks k with stride `inner` (consecutive lanes ->
// consecutive banks); UO outputs per thread share each b[k].
//
// Replaces, for this regime, the same reference code as the other kernels (detail/matrix_times_vector.h:51-127 inside
// the loop nest detail/tensor_times_vector.h:189-324).
#pragma once

#include "numeric.cuh"

namespace ttvb {

struct StreamParams {
  const void* a;
  const void* b;
  void*       c;
  uint64_t outer, nq, inner;
  uint64_t slabs_per_chunk;   // S
  uint64_t chunks;            // ceil(outer / S)
  uint64_t total_bytes;       // bytes of A
  uint32_t stage_bytes;       // shared-memory bytes of one stage (>= S*nq*inner*sizeof(T) + 32, multiple of 128)
  uint32_t accumulate;
};

namespace tma {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void fence_barrier_init()
{ asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async()
{ asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{ asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (1-D TMA): 16-byte aligned addresses, size a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
} // namespace tma

// UO outputs (u, u + step, ...) of one thread: acc_j = sum_k slab[s_j][k][i_j] * b[k], b[k] read once per k for all of them.
// KR partial sums per output (k, k+1, .. go round robin): with UO = 1 -- the stragglers of a chunk, or every output of
// a slab that is alone in its stage and has few more outputs than the CTA has threads -- a single accumulator is one
// dependent FMA chain of n_q shared-memory loads, which 8 warps per SM cannot hide (23 x 529 floats: 4.5 TB/s).
template<class T, int UO, int KR = 1>
__device__ __forceinline__ void stream_outputs(const T* slab0, const T* sb, T* cbase, uint32_t u0, uint32_t step, uint32_t nq,
                                               uint32_t inner, uint32_t M, uint32_t accumulate)
{
  static_assert(KR == 1 || KR == 2 || KR == 4, "partial sums per output");
  T acc[UO][KR];
  const T* base[UO];
#pragma unroll
  for (int j = 0; j < UO; ++j) {
    const uint32_t u = u0 + j * step;
    const uint32_t s = inner == 1 ? u : u / inner;
    const uint32_t i = inner == 1 ? 0 : u - s * inner;
    base[j] = slab0 + (size_t)s * M + i;
#pragma unroll
    for (int r = 0; r < KR; ++r) acc[j][r] = Num<T>::zero();
  }
  uint32_t k = 0;
  for (; k + 4 <= nq; k += 4) {                                    // four k per step: b comes in as one vector when it can
    T bk[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) bk[r] = sb[k + r];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < UO; ++j) acc[j][r % KR] = Num<T>::madd(base[j][(size_t)(k + r) * inner], bk[r], acc[j][r % KR]);
  }
  for (; k < nq; ++k) {
    const T bk = sb[k];
#pragma unroll
    for (int j = 0; j < UO; ++j) acc[j][0] = Num<T>::madd(base[j][(size_t)k * inner], bk, acc[j][0]);
  }
#pragma unroll
  for (int j = 0; j < UO; ++j) {
    T sum = acc[j][0];
#pragma unroll
    for (int r = 1; r < KR; ++r) sum = Num<T>::add(sum, acc[j][r]);
    T* out = cbase + u0 + j * step;
    *out = accumulate ? Num<T>::add(*out, sum) : sum;
  }
}

// Fibers (inner == 1) of EVEN length: lane u reading element k of fiber u hits bank (u*nq + k) % 32, which collides
// for even nq.  Each lane therefore starts its fiber at k = u % nq and wraps around: lane u then reads word
// u*(nq + 1) + t, an odd stride, conflict-free.  b is read per lane at the same rotated index.
template<class T, int UO>
__device__ __forceinline__ void stream_fibers_skewed(const T* slab0, const T* sb, T* cbase, uint32_t u0, uint32_t step, uint32_t nq,
                                                     uint32_t accumulate)
{
  T acc[UO];
  const T* base[UO];
  uint32_t kk[UO];
#pragma unroll
  for (int j = 0; j < UO; ++j) {
    const uint32_t u = u0 + j * step;
    base[j] = slab0 + (size_t)u * nq;
    kk[j] = u % nq;
    acc[j] = Num<T>::zero();
  }
#pragma unroll 4
  for (uint32_t t = 0; t < nq; ++t) {
#pragma unroll
    for (int j = 0; j < UO; ++j) {
      acc[j] = Num<T>::madd(base[j][kk[j]], sb[kk[j]], acc[j]);
      kk[j] = (kk[j] + 1 == nq) ? 0u : kk[j] + 1;
    }
  }
#pragma unroll
  for (int j = 0; j < UO; ++j) {
    T* out = cbase + u0 + j * step;
    *out = accumulate ? Num<T>::add(*out, acc[j]) : acc[j];
  }
}

// MAXT: 256 threads, 2 CTAs per SM for the shared stages of small slabs; up to 1024 threads, 1 CTA per SM when a big
// slab is alone in its stage (the 8 warps of a 256-thread CTA cannot hide the shared-memory latency of 500+ outputs:
// 23 x 529 floats ran at 4.5 TB/s with 30 % of the issue slots busy and 6.5 cycles between two instructions of a warp)
// partial sums per output on the one-output-per-thread path (stages that hold fewer outputs than UO rounds of the CTA)
#ifndef TTVB_STREAM_KR
#define TTVB_STREAM_KR(MAXT) ((MAXT) > 256 ? 4 : 1)
#endif

template<class T, int NS, int UO, bool SKEW = false, int MAXT = 256>
__global__ void __launch_bounds__(MAXT, MAXT == 256 ? 2 : 1)
ttv_stream_kernel(const StreamParams P)
{
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem_raw[];   // the runtime places dynamic shared memory at a 1024-byte aligned offset when it is the only shared allocation
  unsigned char* stages = smem_raw;                                                  // [NS][stage_bytes]
  T*        sb   = reinterpret_cast<T*>(smem_raw + (size_t)NS * P.stage_bytes);      // [nq]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * P.stage_bytes + (((size_t)P.nq * sizeof(T) + 15) & ~(size_t)15));

  const unsigned char* __restrict__ Ab = static_cast<const unsigned char*>(P.a);
  const T* __restrict__ A = static_cast<const T*>(P.a);
  const T* __restrict__ B = static_cast<const T*>(P.b);
  T* __restrict__       C = static_cast<T*>(P.c);

  const uint32_t tid = threadIdx.x;
  const uint32_t nq = (uint32_t)P.nq, inner = (uint32_t)P.inner;
  const uint32_t M = nq * inner;                                   // elements per slab
  const uint64_t tail_begin = P.total_bytes & ~(uint64_t)15;       // bytes of A that bulk copies may touch

  if (tid == 0) {
#pragma unroll
    for (int st = 0; st < NS; ++st) tma::mbar_init(&bars[st], 1);
    tma::fence_barrier_init();
  }
  for (uint32_t j = tid; j < nq; j += blockDim.x) sb[j] = B[j];
  __syncthreads();

  // one elected thread feeds the ring: chunk -> contiguous byte range rounded outwards to 16 bytes
  auto issue = [&](uint64_t chunk, int st) {
    const uint64_t o0 = chunk * P.slabs_per_chunk;
    const uint64_t ns = min(P.slabs_per_chunk, P.outer - o0);
    const uint64_t g0 = o0 * M * sizeof(T), g1 = (o0 + ns) * M * sizeof(T);
    const uint64_t lo = g0 & ~(uint64_t)15;
    uint64_t hi = (g1 + 15) & ~(uint64_t)15;
    if (hi > tail_begin) hi = tail_begin;                          // never read past the end of A
    if (hi > lo) {
      const uint32_t bytes = (uint32_t)(hi - lo);
      tma::mbar_expect_tx(&bars[st], bytes);
      tma::bulk_g2s(stages + (size_t)st * P.stage_bytes, Ab + lo, bytes, &bars[st]);
    } else {
      tma::mbar_arrive(&bars[st]);
    }
  };

  const uint64_t first = blockIdx.x, stride = gridDim.x;
  if (tid == 0) {
#pragma unroll
    for (int st = 0; st < NS; ++st)
      if (first + (uint64_t)st * stride < P.chunks) issue(first + (uint64_t)st * stride, st);
  }

  uint32_t it = 0;
  for (uint64_t chunk = first; chunk < P.chunks; chunk += stride, ++it) {
    const int st = (int)(it % NS);
    const uint32_t parity = (it / NS) & 1u;
    tma::mbar_wait(&bars[st], parity);

    const uint64_t o0 = chunk * P.slabs_per_chunk;
    const uint32_t ns = (uint32_t)min(P.slabs_per_chunk, P.outer - o0);
    const uint64_t g0 = o0 * M * sizeof(T), g1 = g0 + (uint64_t)ns * M * sizeof(T);
    unsigned char* stage = stages + (size_t)st * P.stage_bytes;
    T* slab0 = reinterpret_cast<T*>(stage + (g0 - (g0 & ~(uint64_t)15)));    // element 0 of slab o0 inside the stage

    // the last < 16 bytes of A are not covered by a bulk copy: fetch them with plain loads
    if (g1 > tail_begin) {
      const uint64_t from = max(tail_begin, g0);
      const uint32_t cnt = (uint32_t)((g1 - from) / sizeof(T));
      if (tid < cnt) slab0[(from - g0) / sizeof(T) + tid] = A[from / sizeof(T) + tid];
      __syncthreads();
    }

    const uint32_t outs = ns * inner;                               // outputs of this chunk
    if constexpr (SKEW) {                                           // inner == 1, even n_q (chosen by the launcher)
      for (uint32_t u0 = tid; u0 < outs; u0 += UO * blockDim.x) {
        if (u0 + (UO - 1) * blockDim.x < outs) stream_fibers_skewed<T, UO>(slab0, sb, C + o0, u0, blockDim.x, nq, P.accumulate);
        else
          for (uint32_t u = u0; u < outs; u += blockDim.x) stream_fibers_skewed<T, 1>(slab0, sb, C + o0, u, blockDim.x, nq, P.accumulate);
      }
    } else {
      for (uint32_t u0 = tid; u0 < outs; u0 += UO * blockDim.x) {
        // outputs u0, u0 + NT, ... of this thread; a full set of UO shares every b[k], stragglers go one by one
        if (u0 + (UO - 1) * blockDim.x < outs) stream_outputs<T, UO, (UO == 1 ? 4 : 1)>(slab0, sb, C + o0 * inner, u0, blockDim.x, nq, inner, M, P.accumulate);
        else
          for (uint32_t u = u0; u < outs; u += blockDim.x) stream_outputs<T, 1, TTVB_STREAM_KR(MAXT)>(slab0, sb, C + o0 * inner, u, blockDim.x, nq, inner, M, P.accumulate);
      }
    }

    // everybody is done with this stage: hand it back to the TMA unit for chunk it + NS
    __syncthreads();
    if (tid == 0) {
      const uint64_t next = chunk + (uint64_t)NS * stride;
      if (next < P.chunks) {
        tma::fence_proxy_async();
        issue(next, st);
      }
    }
  }
}

} // namespace ttvb
