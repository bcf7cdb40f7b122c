// colt_kernel.cuh -- COLT: the column GEMV with A staged through shared memory by TMA TENSOR tiles (sm_100a).
//
// BASELINE.json's north_star asks for "TMA bulk tiles for large inner extents"; this is that kernel, built to be measured
// against ttv_col_kernel (kernels.cuh), whose threads load A with plain 16-byte LDG.  A[outer][n_q][inner] is described to
// the TMA unit as a 3-D tensor of 32-bit words (inner*s/4, n_q, outer); one cp.async.bulk.tensor.3d (SASS: UTMALDG) brings a
// box of WT words x KT rows of one slab into a shared-memory stage and signals an mbarrier with the bytes it wrote.  Rows
// and columns beyond the tensor are zero-filled by the hardware, so there is no predication anywhere in the main loop.
//
//   producer   one extra warp; its elected lane walks the CTA's boxes, waits for a stage to be EMPTY, arms the stage's FULL
//              barrier with the box size and issues the copy.  NS stages of WT*KT*4 bytes are in flight per CTA.
//   consumers  256 threads as (ty, tx): tx = WT/4 lanes along inner, each owning 16 bytes (V elements) of every row;
//              ty = 256/tx lanes along n_q take rows ty, ty+TY, ... of a box from shared memory (LDS.128, a warp reads 512
//              contiguous bytes: conflict-free).  After the last box of a column tile the TY partial sums meet in shared
//              memory and lanes ty == 0 store V outputs each.
//
// A work item is one column tile: WT words of `inner` for one slab o, all n_q rows (or one of ksplit partitions of them,
// whose partial sums ttv_reduce_kernel adds up).  CTAs are persistent and stride over the items.  The arithmetic is the reference's gemv_col (detail/matrix_times_vector.h:108-127).
#pragma once

#include <cuda.h>            // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "numeric.cuh"
#include "stream_kernel.cuh" // tma:: mbarrier helpers

namespace ttvb {

struct ColtParams {
  const void* b;
  void*       c;
  uint64_t outer, nq, inner;     // inner in ELEMENTS
  uint64_t itiles, items;        // column tiles per slab, itiles * outer
  uint32_t wt;                   // words (4 bytes) of a row per box: multiple of 4, <= 256
  uint32_t kt;                   // rows per box, <= 256
  uint32_t kboxes;               // boxes per work item: ceil(n_q / kt) / ksplit
  uint32_t ksplit;               // n_q partitions across work items (> 1: partials go to the workspace [ksplit][outer*inner])
  uint32_t tx, ty;               // consumer tile: tx = wt / 4, ty = 256 / tx
  uint32_t stages;
  uint32_t accumulate;
};

namespace tma {
__device__ __forceinline__ void tensor_g2s_3d(void* dst, const CUtensorMap* map, uint32_t c0, uint32_t c1, uint32_t c2, uint64_t* bar)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map)
{ asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory"); }
} // namespace tma

constexpr int kColtConsumers = 256;
constexpr int kColtThreads = kColtConsumers + 32;

template<class T, int V>
__global__ void __launch_bounds__(kColtThreads, 1)
ttv_colt_kernel(const __grid_constant__ CUtensorMap map, const ColtParams P)
{
  pdl_prologue();
  static_assert(sizeof(T) * V == 16, "a consumer lane owns 16 bytes of a row");
  extern __shared__ __align__(128) unsigned char colt_smem[];
  const uint32_t stage_bytes = P.wt * P.kt * 4u;
  unsigned char* stage0 = colt_smem;                                                   // [stages][kt][wt] words
  const uint32_t nb_pad = P.kboxes * P.ksplit * P.kt;                                 // b padded with zeros to whole boxes
  T* sb = reinterpret_cast<T*>(colt_smem + (size_t)P.stages * stage_bytes);           // [nb_pad]
  T* red = sb + (((size_t)nb_pad * sizeof(T) + 15) / 16 * 16) / sizeof(T);           // [256][V]
  uint64_t* full = reinterpret_cast<uint64_t*>(red + (size_t)kColtConsumers * V);    // [stages]
  uint64_t* empty = full + P.stages;                                                  // [stages]

  const uint32_t tid = threadIdx.x;
  const T* __restrict__ B = static_cast<const T*>(P.b);
  T* __restrict__ C = static_cast<T*>(P.c);

  if (tid == 0) {
    for (uint32_t s = 0; s < P.stages; ++s) { tma::mbar_init(&full[s], 1); tma::mbar_init(&empty[s], kColtConsumers / 32); }
    tma::fence_barrier_init();
  }
  for (uint32_t j = tid; j < nb_pad; j += blockDim.x) sb[j] = j < (uint32_t)P.nq ? B[j] : Num<T>::zero();
  __syncthreads();

  if (tid >= kColtConsumers) {
    // ---- producer warp ----
    if (tid == kColtConsumers) {
      tma::prefetch_map(&map);
      uint32_t stage = 0, phase = 0;
      for (uint64_t item = blockIdx.x; item < P.items; item += gridDim.x) {
        const uint32_t it = (uint32_t)(item % P.itiles);
        const uint64_t r  = item / P.itiles;
        const uint32_t ks = (uint32_t)(r % P.ksplit);
        const uint32_t o  = (uint32_t)(r / P.ksplit);
        for (uint32_t kb = 0; kb < P.kboxes; ++kb) {
          const uint32_t k0 = (ks * P.kboxes + kb) * P.kt;                            // (a box wholly past n_q arrives as zeros)
          tma::mbar_wait(&empty[stage], phase ^ 1u);                                  // the consumers are done with this stage
          tma::mbar_expect_tx(&full[stage], stage_bytes);
          tma::tensor_g2s_3d(stage0 + (size_t)stage * stage_bytes, &map, it * P.wt, k0, o, &full[stage]);
          if (++stage == P.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    return;
  }

  // ---- consumers ----
  const uint32_t tx = tid % P.tx;
  const uint32_t ty = tid / P.tx;                       // < P.ty (tx * ty == 256)
  const uint32_t rows_per_lane = P.kt / P.ty;           // kt is a multiple of ty
  uint32_t stage = 0, phase = 0;
  for (uint64_t item = blockIdx.x; item < P.items; item += gridDim.x) {
    const uint64_t it = item % P.itiles;
    const uint64_t rr = item / P.itiles;
    const uint32_t ks = (uint32_t)(rr % P.ksplit);
    const uint64_t o  = rr / P.ksplit;
    T acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = Num<T>::zero();
    for (uint32_t kb = 0; kb < P.kboxes; ++kb) {
      tma::mbar_wait(&full[stage], phase);                                            // the box has landed
      const unsigned char* box = stage0 + (size_t)stage * stage_bytes;
      const uint32_t k0 = (ks * P.kboxes + kb) * P.kt;                                // rows past n_q: zeros in the box, zeros in sb
#pragma unroll 4
      for (uint32_t r = 0; r < rows_per_lane; ++r) {
        const uint32_t k = ty + r * P.ty;
        const Vec<T, V> v = *reinterpret_cast<const Vec<T, V>*>(box + ((size_t)k * P.wt + tx * 4u) * 4u);
        const T bb = sb[k0 + k];
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = Num<T>::madd(v.e[j], bb, acc[j]);
      }
      __syncwarp();
      if ((tid & 31u) == 0) tma::mbar_arrive(&empty[stage]);                          // one arrival per consumer warp
      if (++stage == P.stages) { stage = 0; phase ^= 1u; }
    }
    // the ty partial sums of this column tile meet in shared memory (consumers only: named barrier 1)
    T* mine = red + (size_t)tid * V;
    if (P.ty > 1) {
#pragma unroll
      for (int j = 0; j < V; ++j) mine[j] = acc[j];
      asm volatile("bar.sync 1, %0;" ::"n"(kColtConsumers) : "memory");
      if (ty == 0) {
        for (uint32_t y = 1; y < P.ty; ++y) {
          const T* other = mine + (size_t)y * P.tx * V;
#pragma unroll
          for (int j = 0; j < V; ++j) acc[j] = Num<T>::add(acc[j], other[j]);
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kColtConsumers) : "memory");
    }
    const uint64_t i0 = (it * P.wt + tx * 4u) * 4u / sizeof(T);                       // first element of this lane's 16 bytes
    if (ty == 0 && i0 < P.inner) {
      Vec<T, V>* out = reinterpret_cast<Vec<T, V>*>(C + (P.ksplit > 1 ? (uint64_t)ks * P.outer * P.inner : 0) + o * P.inner + i0);
      Vec<T, V> val;
      if (P.accumulate) {
        const Vec<T, V> old = *out;
#pragma unroll
        for (int j = 0; j < V; ++j) val.e[j] = Num<T>::add(old.e[j], acc[j]);
      } else {
#pragma unroll
        for (int j = 0; j < V; ++j) val.e[j] = acc[j];
      }
      *out = val;
    }
  }
}

} // namespace ttvb
