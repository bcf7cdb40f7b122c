// scatter_kernel.cuh -- the n_q-split product of the multi-GPU path FUSED with its exchange step.
//
// When mode q is the slowest mode of the layout, G GPUs each hold a contiguous range of the contraction rows and compute
// a full-size PARTIAL of C; the partials have to be summed ("one ncclReduce").  Here the exchange rides on the
// kernel's own stores: the flat index space of C is cut into G blocks of `blk` elements, block j belongs to GPU j, and
// every GPU writes the piece of its partial that falls into block j straight into GPU j's memory -- slot `rank` of
// j's workspace [G][blk], mapped into this process over NVLink / NVSwitch (peer pointers from symmetric memory).  The
// 16-byte stores leave while the kernel is still streaming A, so the transfer (|C|·(G-1)/G bytes per GPU) hides behind
// the HBM-bound main loop.  After a cross-GPU barrier each GPU sums its G slots in rank order (ttv_reduce_kernel:
// deterministic) and holds its block of C: a reduce-scatter, the same distributed form the free-split products leave C in.
//
// The arithmetic is the column GEMV of kernels.cuh (reference gemv_col, detail/matrix_times_vector.h:108-127) with a
// whole CTA along inner and eight k-steps in flight.
#pragma once

#include "kernels.cuh"

namespace ttvb {

constexpr int kMaxPeers = 16;

struct ScatterParams {
  const void* a;
  const void* b;
  void*       peer[kMaxPeers];   // peer[j]: GPU j's workspace [world][blk], addressable from this GPU
  uint64_t outer, nq, inner;
  uint64_t blk;                  // elements of C's flat index space per GPU (multiple of the vector width)
  uint64_t itiles, tiles;
  uint32_t world, rank;
  uint32_t kb;                   // elements of b per shared-memory chunk
  uint32_t stream;
};

// the product + scatter part shared by both kernels below: every tile's partial sums leave as 16-byte stores into the owner's slot
template<class T, int V>
__device__ __forceinline__ void scatter_tiles(const ScatterParams& P, T* sb)
{
  constexpr int KU = 8;
  const T* __restrict__ A = static_cast<const T*>(P.a);
  const T* __restrict__ B = static_cast<const T*>(P.b);
  const uint32_t tid = threadIdx.x;
  const bool stream = P.stream != 0;
  const bool b_resident = P.nq <= P.kb;

  if (b_resident) {
    for (uint32_t j = tid; j < (uint32_t)P.nq; j += blockDim.x) sb[j] = B[j];
    __syncthreads();
  }

  for (uint64_t tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
    const uint64_t it = tile % P.itiles;
    const uint64_t o  = tile / P.itiles;
    const uint64_t i0 = (it * blockDim.x + tid) * V;
    const int nvalid = i0 < P.inner ? 1 : 0;

    T acc[1][V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[0][j] = Num<T>::zero();

    for (uint64_t k0 = 0; k0 < P.nq; k0 += P.kb) {
      const uint32_t kn = (uint32_t)min((uint64_t)P.kb, P.nq - k0);
      if (!b_resident) {
        __syncthreads();
        for (uint32_t j = tid; j < kn; j += blockDim.x) sb[j] = B[k0 + j];
        __syncthreads();
      }
      if (nvalid) {
        const T* ap = A + (o * P.nq + k0) * P.inner + i0;
        uint32_t k = 0;
        for (; k + (KU - 1) < kn; k += KU, ap += KU * P.inner)
          col_batch<T, V, 1, KU, false>(acc, ap, 0, P.inner, sb, k, 1u, kn, 1, stream);
        for (; k < kn; k += KU, ap += KU * P.inner)
          col_batch<T, V, 1, KU, true>(acc, ap, 0, P.inner, sb, k, 1u, kn, 1, stream);
      }
    }

    if (nvalid) {
      // the owner of this vector of C and the slot this GPU writes there
      const uint64_t f = o * P.inner + i0;
      const uint64_t j = f / P.blk;
      T* dst = static_cast<T*>(P.peer[j]) + (uint64_t)P.rank * P.blk + (f - j * P.blk);
      Vec<T, V> val;
#pragma unroll
      for (int e = 0; e < V; ++e) val.e[e] = acc[0][e];
      *reinterpret_cast<Vec<T, V>*>(dst) = val;
    }
  }
}

template<class T, int V>
__global__ void __launch_bounds__(256, 3)
ttv_col_scatter_kernel(const ScatterParams P)
{
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  scatter_tiles<T, V>(P, reinterpret_cast<T*>(smem_raw));            // [kb]
}

// ------------------------------------------------------------------------------------------------------------------
// The WHOLE exchange in one kernel: product + scatter (above), a barrier across the GPUs, and the sum of the received slots.
//
//   1  every CTA computes its tiles and stores the partial sums into the owners' slots over NVLink (scatter_tiles);
//   2  it fences to system scope and counts itself in.  All but the LAST `reducers` CTAs to arrive are done and exit -- the
//      grid is the same oversubscribed one the plain kernels use (64 CTAs per SM queued), nothing has to be co-resident.
//      The very last arrival publishes "GPU `rank` has delivered round `token`" by writing the token into flag[rank] of
//      EVERY GPU's flag array (peer stores), its own included;
//   3  the last `reducers` arrivals (one per SM by default) wait until all `world` flags of their OWN array show the token
//      (ld.acquire.sys): all partials for this GPU's block have landed.  They hold at most a third of the CTA slots, so the
//      CTAs that still have to arrive always find room: no deadlock;
//   4  they sum the `world` slots of this GPU's workspace in rank order into its block of C (deterministic).
//
// Tokens only grow, the two workspace halves alternate (see sharded.PeerExchange), so nothing is ever reset across GPUs; the
// local arrival counter is reset by the CTA that closes it.  A wait that lasts longer than `timeout_ns` sets *error and gives
// up (a peer that never launched must not hang the GPU).
// This replaces scatter kernel + library barrier kernel + reduce kernel by one launch (SURVEY 8e "one ncclReduce").
// (A first version ran a persistent grid of resident CTAs that all waited: 3 % slower on 2 GPUs than the three-launch form,
// the 8 192 tiles of a 34 GB slab over 444 CTAs leave the last wave half empty.)
// ------------------------------------------------------------------------------------------------------------------
struct ExchangeParams {
  ScatterParams S;
  void*     c;                          // this GPU's block of C, n_block elements
  uint64_t  n_block;
  uint32_t* flags_peer[kMaxPeers];      // GPU j's flag array [kMaxPeers], addressable from this GPU
  unsigned long long* counter;          // local: CTAs of this launch that have delivered
  uint32_t* error;                      // local: set to 1 when a wait timed out
  unsigned long long timeout_ns;
  uint32_t  token;                      // round number, grows by one per exchange of the group
  uint32_t  accumulate;
  uint32_t  reducers;                   // how many of the last CTAs to arrive stay for the barrier and the slot sum
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v)
{
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// coherent load that bypasses L1 (ld.global.cg): the slots are written by OTHER GPUs while this kernel runs, so the
// read-only path (ld.global.nc) is not allowed for them
template<class T, int V>
__device__ __forceinline__ Vec<T, V> load_cg(const T* p)
{
  Vec<T, V> v;
  constexpr int BYTES = (int)sizeof(T) * V;
  if constexpr (BYTES == 16) {
    uint32_t x, y, z, w;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "l"(p) : "memory");
    uint32_t* d = reinterpret_cast<uint32_t*>(&v); d[0] = x; d[1] = y; d[2] = z; d[3] = w;
  } else if constexpr (BYTES == 8) {
    uint32_t x, y;
    asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "l"(p) : "memory");
    uint32_t* d = reinterpret_cast<uint32_t*>(&v); d[0] = x; d[1] = y;
  } else {
    static_assert(BYTES == 4, "vector width");
    uint32_t x;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(x) : "l"(p) : "memory");
    *reinterpret_cast<uint32_t*>(&v) = x;
  }
  return v;
}

__device__ __forceinline__ unsigned long long global_timer_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template<class T, int V>
__global__ void __launch_bounds__(256, 3)
ttv_col_exchange_kernel(const ExchangeParams E)
{
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const ScatterParams& P = E.S;
  scatter_tiles<T, V>(P, reinterpret_cast<T*>(smem_raw));

  // 2: deliver.  The CTA orders its peer stores before its arrival; the closing CTA orders all arrivals before the flags
  // (fence - atomic ... atomic - fence).
  __shared__ unsigned long long s_arrival;
  __syncthreads();                                                     // every thread's peer stores are issued ...
  if (threadIdx.x == 0) {
    __threadfence_system();                                            // ... and ordered, cumulatively, before the arrival (the
    s_arrival = atomicAdd(E.counter, 1ull);                            // pattern of a grid-wide barrier: bar.sync, fence, atomic)
  }
  __syncthreads();
  const unsigned long long arrival = s_arrival;                        // 0 .. gridDim.x - 1
  const unsigned long long R = min((unsigned long long)max(E.reducers, 1u), (unsigned long long)gridDim.x);
  if (arrival + R < (unsigned long long)gridDim.x) return;             // not among the last R: done
  const uint64_t me = arrival - ((unsigned long long)gridDim.x - R);   // reducer index 0 .. R-1
  if (arrival + 1 == (unsigned long long)gridDim.x && threadIdx.x == 0) {
    *E.counter = 0ull;                                                 // ready for the next launch on this stream
    __threadfence_system();
    for (uint32_t j = 0; j < P.world; ++j) st_release_sys(E.flags_peer[j] + P.rank, E.token);
  }
  // 3: wait for every GPU's delivery into OUR workspace
  if (threadIdx.x < P.world) {
    const uint32_t* mine = E.flags_peer[P.rank] + threadIdx.x;
    const unsigned long long t0 = global_timer_ns();
    while ((int32_t)(ld_acquire_sys(mine) - E.token) < 0) {
      if (global_timer_ns() - t0 > E.timeout_ns) { *E.error = 1u; break; }
      __nanosleep(200);
    }
  }
  __syncthreads();

  // 4: sum the slots of our block in rank order
  const T* ws = static_cast<const T*>(P.peer[P.rank]);
  T* __restrict__ C = static_cast<T*>(E.c);
  const uint64_t nvec = E.n_block / V;                                 // (n_block is a multiple of V except for the tail below)
  for (uint64_t j = me * blockDim.x + threadIdx.x; j < nvec; j += R * blockDim.x) {
    Vec<T, V> sum;
    if (E.accumulate) sum = *reinterpret_cast<const Vec<T, V>*>(C + j * V);
    else {
#pragma unroll
      for (int e = 0; e < V; ++e) sum.e[e] = Num<T>::zero();
    }
    for (uint32_t r = 0; r < P.world; ++r) {
      const Vec<T, V> x = load_cg<T, V>(ws + (uint64_t)r * P.blk + j * V);          // coherent, L1 bypassed: the data came in over NVLink
#pragma unroll
      for (int e = 0; e < V; ++e) sum.e[e] = Num<T>::add(sum.e[e], x.e[e]);
    }
    *reinterpret_cast<Vec<T, V>*>(C + j * V) = sum;
  }
  for (uint64_t j = nvec * V + me * blockDim.x + threadIdx.x; j < E.n_block; j += R * blockDim.x) {
    T sum = E.accumulate ? C[j] : Num<T>::zero();
    for (uint32_t r = 0; r < P.world; ++r) sum = Num<T>::add(sum, load_cg<T, 1>(ws + (uint64_t)r * P.blk + j).e[0]);
    C[j] = sum;
  }
}

} // namespace ttvb
