// launch.cu -- instantiates the kernels of kernels.cuh and launches them.
//
// Compiled several times by ttv_b200/build.py (in parallel):
//   without TTVB_DTYPE   -> the dtype-independent part: launch_view / launch_fill / counters
//   with -DTTVB_DTYPE=k  -> the kernels of element type k (enum ttv_b200_dtype) and their dispatcher
#include "launch.h"
#include "kernels.cuh"
#include "stream_kernel.cuh"
#include "colx_kernel.cuh"
#include "colr_kernel.cuh"
#include "dotf_kernel.cuh"
#include "strided_kernel.cuh"
#include "scatter_kernel.cuh"
#include "colt_kernel.cuh"
#include "streamk_kernel.cuh"
#include "dotp_kernel.cuh"
#include "colf_kernel.cuh"

#include <algorithm>
#include <cstdlib>
#include <atomic>
#include <utility>

namespace ttvb {

// one dispatcher per element type, defined in the TTVB_DTYPE translation units
using tile_fn_t   = cudaError_t (*)(const TileParams&, const Launch&, cudaStream_t);
using reduce_fn_t = cudaError_t (*)(const void*, void*, uint64_t, uint32_t, bool, uint64_t, int, cudaStream_t);
using scatter_fn_t = cudaError_t (*)(const ScatterParams&, int, uint64_t, uint64_t, cudaStream_t);
using exchange_fn_t = cudaError_t (*)(const ExchangeParams&, int, uint32_t, int, uint64_t, cudaStream_t);
using fill_fn_t   = cudaError_t (*)(void*, uint64_t, uint64_t, uint64_t, int, cudaStream_t);
using stream_fn_t = cudaError_t (*)(const StreamParams&, const Launch&, cudaStream_t);
using dotf_fn_t   = cudaError_t (*)(const DotfParams&, const Launch&, cudaStream_t);
using strided_fn_t = cudaError_t (*)(const StridedParams&, int, cudaStream_t);
using streamk_fn_t = cudaError_t (*)(const StreamkParams&, const Launch&, cudaStream_t);
using dotp_fn_t   = cudaError_t (*)(const DotpParams&, const Launch&, cudaStream_t);
using colf_fn_t   = cudaError_t (*)(const ColfParams&, const Launch&, cudaStream_t);
using colf_tiny_fn_t = cudaError_t (*)(const ColfTinyParams&, const Launch&, cudaStream_t);
using colt_fn_t   = cudaError_t (*)(const CUtensorMap&, const ColtParams&, const Launch&, cudaStream_t);

#define TTVB_DECLARE(k)                                                                                          \
  cudaError_t tile_dtype_##k(const TileParams&, const Launch&, cudaStream_t);                                    \
  cudaError_t reduce_dtype_##k(const void*, void*, uint64_t, uint32_t, bool, uint64_t, int, cudaStream_t);       \
  cudaError_t scatter_dtype_##k(const ScatterParams&, int, uint64_t, uint64_t, cudaStream_t);                    \
  cudaError_t exchange_dtype_##k(const ExchangeParams&, int, uint32_t, int, uint64_t, cudaStream_t);             \
  cudaError_t fill_dtype_##k(void*, uint64_t, uint64_t, uint64_t, int, cudaStream_t);                            \
  cudaError_t stream_dtype_##k(const StreamParams&, const Launch&, cudaStream_t);                               \
  cudaError_t dotf_dtype_##k(const DotfParams&, const Launch&, cudaStream_t);                                   \
  cudaError_t strided_dtype_##k(const StridedParams&, int, cudaStream_t);                                         \
  cudaError_t colt_dtype_##k(const CUtensorMap&, const ColtParams&, const Launch&, cudaStream_t);                \
  cudaError_t streamk_dtype_##k(const StreamkParams&, const Launch&, cudaStream_t);                             \
  cudaError_t dotp_dtype_##k(const DotpParams&, const Launch&, cudaStream_t);                                   \
  cudaError_t colf_dtype_##k(const ColfParams&, const Launch&, cudaStream_t);                                   \
  cudaError_t colf_tiny_dtype_##k(const ColfTinyParams&, const Launch&, cudaStream_t);
TTVB_DECLARE(0) TTVB_DECLARE(1) TTVB_DECLARE(2) TTVB_DECLARE(3) TTVB_DECLARE(4) TTVB_DECLARE(5)
#undef TTVB_DECLARE

void count_launch();

// Every kernel goes out through cudaLaunchKernelEx.  With TTV_B200_PDL (default on) the launch carries the
// programmatic-stream-serialization attribute: the kernel may start placing CTAs while the tail of the previous kernel of
// the stream is still running and waits inside (pdl_prologue, numeric.cuh) until that kernel has completed -- stream
// semantics unchanged, launch latency and ramp hidden (an 80 us product on a 512 MiB tensor: ~4 % of its time).
static inline bool pdl_enabled()
{
  const char* e = std::getenv("TTV_B200_PDL");
  return (e && *e) ? std::atoi(e) != 0 : true;
}

template<class... KArgs, class... Args>
static inline cudaError_t launch_k(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream, Args&&... args)
{
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(block, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1u : 0u;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
  count_launch();
  return e != cudaSuccess ? e : cudaGetLastError();
}

#ifndef TTVB_DTYPE
// ===================================================================================================================
static std::atomic<uint64_t> g_launches{0};
uint64_t launch_count() { return g_launches.load(); }
void count_launch() { g_launches.fetch_add(1); }

static const tile_fn_t   k_tile[]   = {tile_dtype_0, tile_dtype_1, tile_dtype_2, tile_dtype_3, tile_dtype_4, tile_dtype_5};
static const reduce_fn_t k_reduce[] = {reduce_dtype_0, reduce_dtype_1, reduce_dtype_2, reduce_dtype_3, reduce_dtype_4, reduce_dtype_5};
static const fill_fn_t   k_fill[]   = {fill_dtype_0, fill_dtype_1, fill_dtype_2, fill_dtype_3, fill_dtype_4, fill_dtype_5};
static const stream_fn_t k_stream[] = {stream_dtype_0, stream_dtype_1, stream_dtype_2, stream_dtype_3, stream_dtype_4, stream_dtype_5};
static const dotf_fn_t   k_dotf[]   = {dotf_dtype_0, dotf_dtype_1, dotf_dtype_2, dotf_dtype_3, dotf_dtype_4, dotf_dtype_5};
static const scatter_fn_t k_scatter[] = {scatter_dtype_0, scatter_dtype_1, scatter_dtype_2, scatter_dtype_3, scatter_dtype_4, scatter_dtype_5};
static const exchange_fn_t k_exchange[] = {exchange_dtype_0, exchange_dtype_1, exchange_dtype_2, exchange_dtype_3, exchange_dtype_4, exchange_dtype_5};
static const strided_fn_t k_strided[] = {strided_dtype_0, strided_dtype_1, strided_dtype_2, strided_dtype_3, strided_dtype_4, strided_dtype_5};
static const streamk_fn_t k_streamk[] = {streamk_dtype_0, streamk_dtype_1, streamk_dtype_2, streamk_dtype_3, streamk_dtype_4, streamk_dtype_5};
static const dotp_fn_t   k_dotp[]   = {dotp_dtype_0, dotp_dtype_1, dotp_dtype_2, dotp_dtype_3, dotp_dtype_4, dotp_dtype_5};
static const colf_fn_t   k_colf[]   = {colf_dtype_0, colf_dtype_1, colf_dtype_2, colf_dtype_3, colf_dtype_4, colf_dtype_5};
static const colf_tiny_fn_t k_colf_tiny[] = {colf_tiny_dtype_0, colf_tiny_dtype_1, colf_tiny_dtype_2, colf_tiny_dtype_3, colf_tiny_dtype_4, colf_tiny_dtype_5};
static const colt_fn_t   k_colt[]   = {colt_dtype_0, colt_dtype_1, colt_dtype_2, colt_dtype_3, colt_dtype_4, colt_dtype_5};

// cuTensorMapEncodeTiled lives in the driver library; the runtime hands out its address, so nothing links against libcuda
using encode_fn_t = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_fn_t tensor_map_encoder()
{
  static encode_fn_t fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      p = nullptr;
    }
    return reinterpret_cast<encode_fn_t>(p);
  }();
  return fn;
}

// COLT: A as a 3-D tensor of 32-bit words (inner*s/4, n_q, outer), boxes of wt words x kt rows of one slab
static cudaError_t launch_colt(int dtype, const View& v, const Launch& l, const void* a, const void* b, void* c, void* workspace,
                               bool accumulate, cudaStream_t stream)
{
  encode_fn_t encode = tensor_map_encoder();
  if (!encode) return cudaErrorNotSupported;
  const uint64_t s = (uint64_t)dtype_size(dtype);
  CUtensorMap map;
  const cuuint64_t gdim[3] = {v.inner * s / 4, v.nq, v.outer};
  const cuuint64_t gstride[2] = {v.inner * s, v.nq * v.inner * s};                 // bytes between rows / between slabs
  const cuuint32_t box[3] = {l.wt, l.kt, 1};
  const cuuint32_t estride[3] = {1, 1, 1};
  if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void*>(a), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return cudaErrorInvalidValue;
  ColtParams P;
  P.b = b; P.c = l.ksplit > 1 ? workspace : c;
  P.outer = v.outer; P.nq = v.nq; P.inner = v.inner;
  P.itiles = l.itiles; P.items = l.tiles;
  P.wt = l.wt; P.kt = l.kt; P.kboxes = l.kboxes;
  P.tx = l.tx; P.ty = l.ty; P.stages = l.stages; P.ksplit = l.ksplit;
  P.accumulate = (accumulate && l.ksplit == 1) ? 1u : 0u;
  return k_colt[dtype](map, P, l, stream);
}

cudaError_t launch_strided(int dtype, const View& v, const void* a, const void* b, void* c, bool accumulate, int sm_count,
                           cudaStream_t stream)
{
  if (dtype < 0 || dtype >= TTV_B200_DTYPE_COUNT || !v.strided || v.nfree > (uint32_t)kMaxFree) return cudaErrorInvalidValue;
  StridedParams S;
  S.a = a; S.b = b; S.c = c;
  S.nq = v.nq; S.wq = v.wq;
  S.total = 1;
  for (uint32_t d = 0; d < (uint32_t)kMaxFree; ++d) {
    S.n[d] = d < v.nfree ? v.fn[d] : 1; S.wa[d] = d < v.nfree ? v.fwa[d] : 0; S.wc[d] = d < v.nfree ? v.fwc[d] : 0;
    S.total *= S.n[d];
  }
  S.nfree = v.nfree;
  S.accumulate = accumulate ? 1u : 0u;
  return k_strided[dtype](S, sm_count, stream);
}

cudaError_t launch_view(int dtype, const View& v, const Launch& l, const void* a, const void* b, void* c,
                        void* workspace, bool accumulate, int sm_count, cudaStream_t stream)
{
  if (dtype < 0 || dtype >= TTV_B200_DTYPE_COUNT) return cudaErrorInvalidValue;
  if (l.kernel == TTV_B200_KERNEL_STREAM) {
    StreamParams S;
    S.a = a; S.b = b; S.c = c;
    S.outer = v.outer; S.nq = v.nq; S.inner = v.inner;
    S.slabs_per_chunk = l.slabs_per_chunk; S.chunks = l.chunks;
    S.total_bytes = v.outer * v.nq * v.inner * (uint64_t)dtype_size(dtype);
    S.stage_bytes = l.stage_bytes;
    S.accumulate = accumulate ? 1u : 0u;
    return k_stream[dtype](S, l, stream);
  }
  if (l.kernel == TTV_B200_KERNEL_STREAMK) {
    StreamkParams K;
    const uint64_t es = (uint64_t)dtype_size(dtype);
    K.a = a; K.b = b; K.c = l.ksplit > 1 ? workspace : c;
    K.outer = v.outer; K.nq = v.nq; K.inner = v.inner;
    K.kchunk = l.kchunk;
    K.total_bytes_a = v.outer * v.nq * v.inner * es; K.total_bytes_b = v.nq * es;
    K.ksplit = l.ksplit; K.rows_per_stage = (uint32_t)l.slabs_per_chunk;
    K.a_stage_bytes = l.stage_bytes; K.b_stage_bytes = l.b_stage_bytes;
    K.accumulate = (accumulate && l.ksplit == 1) ? 1u : 0u;
    cudaError_t e = k_streamk[dtype](K, l, stream);
    if (e != cudaSuccess || l.ksplit <= 1) return e;
    return k_reduce[dtype](workspace, c, v.outer * v.inner, l.ksplit, accumulate, v.outer * v.inner, sm_count, stream);
  }
  if (l.kernel == TTV_B200_KERNEL_COLT) {
    cudaError_t e = launch_colt(dtype, v, l, a, b, c, workspace, accumulate, stream);
    if (e != cudaSuccess || l.ksplit <= 1) return e;
    return k_reduce[dtype](workspace, c, v.outer * v.inner, l.ksplit, accumulate, v.outer * v.inner, sm_count, stream);
  }
  if (l.kernel == TTV_B200_KERNEL_COLF && l.tiny) {
    ColfTinyParams Y;
    Y.a = a; Y.b = b; Y.c = c; Y.outer = v.outer; Y.G = l.ty; Y.accumulate = accumulate ? 1u : 0u;
    return k_colf_tiny[dtype](Y, l, stream);
  }
  if (l.kernel == TTV_B200_KERNEL_COLF) {
    ColfParams F;
    F.a = a; F.b = b; F.c = l.ksplit > 1 ? workspace : c;
    F.outer = v.outer; F.nq = v.nq; F.inner = v.inner;
    F.srchunk = l.kchunk / l.to;                                     // l.to carries R, the rows of a super-row
    F.ksplit = l.ksplit; F.R = l.to; F.L = l.tx; F.TY = l.ty; F.SW = l.nu;
    F.accumulate = (accumulate && l.ksplit == 1) ? 1u : 0u;
    cudaError_t e = k_colf[dtype](F, l, stream);
    if (e != cudaSuccess || l.ksplit <= 1) return e;
    return k_reduce[dtype](workspace, c, v.outer * v.inner, l.ksplit, accumulate, v.outer * v.inner, sm_count, stream);
  }
  if (l.kernel == TTV_B200_KERNEL_DOTP) {
    DotpParams D;
    D.a = a; D.b = b; D.c = c;
    D.outer = v.outer; D.nvec = v.outer * 2 * (uint64_t)dtype_size(dtype) / 16; D.tiles = l.tiles;
    D.accumulate = accumulate ? 1u : 0u;
    return k_dotp[dtype](D, l, stream);
  }
  if (l.kernel == TTV_B200_KERNEL_DOTF) {
    DotfParams D;
    D.a = a; D.b = b; D.c = c;
    D.outer = v.outer; D.chunks = l.chunks;
    D.nq = (uint32_t)v.nq; D.nv = (uint32_t)(v.nq / (uint64_t)l.vec); D.fw = (uint32_t)l.slabs_per_chunk;
    D.accumulate = accumulate ? 1u : 0u;
    D.lpf = 1;
    while (D.lpf * 2 * (D.fw < 32 ? D.fw : 32) <= 32 && D.lpf * 2 <= D.nv) D.lpf *= 2;
    return k_dotf[dtype](D, l, stream);
  }
  TileParams P;
  P.a = a; P.b = b;
  P.c = l.ksplit > 1 ? workspace : c;
  P.outer = v.outer; P.nq = v.nq; P.inner = v.inner;
  P.kchunk = l.kchunk;
  P.itiles = l.itiles; P.otiles = l.otiles; P.tiles = l.tiles;
  P.a_ustride = l.a_ustride; P.c_ustride = l.c_ustride;
  P.tx = l.tx; P.ty = l.ty; P.to = l.to;
  P.ksplit = l.ksplit; P.kb = l.kb;
  P.accumulate = accumulate ? 1u : 0u;
  P.udir = l.udir; P.stream = l.stream;
  if (l.kernel == TTV_B200_KERNEL_COLX) P.c_ustride = l.wcols;

  cudaError_t e = k_tile[dtype](P, l, stream);
  if (e != cudaSuccess || l.ksplit <= 1) return e;
  return k_reduce[dtype](workspace, c, v.outer * v.inner, l.ksplit, accumulate, v.outer * v.inner, sm_count, stream);
}

// Fused n_q-split product + exchange (scatter_kernel.cuh): this GPU's partial of C goes block by block into the peers'
// workspaces.  vec = elements per 16-byte vector usable for (inner, alignments, blk).
cudaError_t launch_scatter(int dtype, const View& v, const void* a, const void* b, void* const* peers, uint32_t world, uint32_t rank,
                           uint64_t blk, int vec, int sm_count, cudaStream_t stream)
{
  if (dtype < 0 || dtype >= TTV_B200_DTYPE_COUNT || world == 0 || world > (uint32_t)kMaxPeers || rank >= world) return cudaErrorInvalidValue;
  ScatterParams S;
  S.a = a; S.b = b;
  for (uint32_t j = 0; j < (uint32_t)kMaxPeers; ++j) S.peer[j] = j < world ? peers[j] : nullptr;
  S.outer = v.outer; S.nq = v.nq; S.inner = v.inner;
  S.blk = blk;
  S.itiles = (v.inner / (uint64_t)vec + 255) / 256;
  S.tiles = S.itiles * v.outer;
  S.world = world; S.rank = rank;
  const uint64_t s = (uint64_t)dtype_size(dtype);
  S.kb = (uint32_t)std::min<uint64_t>(v.nq, 16384 / s);
  S.stream = s >= 16 ? 1u : 0u;
  const uint64_t ctas = std::max<uint64_t>(1, std::min<uint64_t>(S.tiles, (uint64_t)sm_count * 64));
  return k_scatter[dtype](S, vec, ctas, (uint64_t)S.kb * s, stream);
}

cudaError_t launch_exchange(int dtype, const View& v, const void* a, const void* b, void* const* peers, void* const* flags,
                            uint32_t world, uint32_t rank, uint64_t blk, int vec, void* c_block, uint64_t n_block, uint32_t token,
                            void* counter, void* error, bool accumulate, uint64_t timeout_ns, uint32_t max_ctas, int sm_count,
                            cudaStream_t stream)
{
  if (dtype < 0 || dtype >= TTV_B200_DTYPE_COUNT || world == 0 || world > (uint32_t)kMaxPeers || rank >= world) return cudaErrorInvalidValue;
  ExchangeParams E;
  ScatterParams& S = E.S;
  S.a = a; S.b = b;
  for (uint32_t j = 0; j < (uint32_t)kMaxPeers; ++j) {
    S.peer[j] = j < world ? peers[j] : nullptr;
    E.flags_peer[j] = j < world ? static_cast<uint32_t*>(flags[j]) : nullptr;
  }
  S.outer = v.outer; S.nq = v.nq; S.inner = v.inner;
  S.blk = blk;
  S.itiles = (v.inner / (uint64_t)vec + 255) / 256;
  S.tiles = S.itiles * v.outer;
  S.world = world; S.rank = rank;
  const uint64_t s = (uint64_t)dtype_size(dtype);
  S.kb = (uint32_t)std::min<uint64_t>(v.nq, 16384 / s);
  S.stream = s >= 16 ? 1u : 0u;
  E.c = c_block; E.n_block = n_block;
  E.counter = static_cast<unsigned long long*>(counter);
  E.error = static_cast<uint32_t*>(error);
  E.timeout_ns = timeout_ns;
  E.token = token;
  E.accumulate = accumulate ? 1u : 0u;
  return k_exchange[dtype](E, vec, max_ctas, sm_count, (uint64_t)S.kb * s, stream);
}

cudaError_t launch_reduce_slots(int dtype, const void* ws, void* c, uint64_t n, uint64_t stride, uint32_t slots, bool accumulate,
                                int sm_count, cudaStream_t stream)
{
  if (dtype < 0 || dtype >= TTV_B200_DTYPE_COUNT) return cudaErrorInvalidValue;
  if (n == 0) return cudaSuccess;
  return k_reduce[dtype](ws, c, n, slots, accumulate, stride, sm_count, stream);
}

cudaError_t launch_fill(int dtype, void* x, uint64_t first, uint64_t count, uint64_t seed, int sm_count, cudaStream_t stream)
{
  if (dtype < 0 || dtype >= TTV_B200_DTYPE_COUNT) return cudaErrorInvalidValue;
  if (count == 0) return cudaSuccess;
  return k_fill[dtype](x, first, count, seed, sm_count, stream);
}

#else
// ===================================================================================================================
#if   TTVB_DTYPE == 0
using elem_t = float;              constexpr int kVmax = 4;
#elif TTVB_DTYPE == 1
using elem_t = double;             constexpr int kVmax = 2;
#elif TTVB_DTYPE == 2
using elem_t = cf32;               constexpr int kVmax = 2;
#elif TTVB_DTYPE == 3
using elem_t = cf64;               constexpr int kVmax = 1;
#elif TTVB_DTYPE == 4
using elem_t = uint32_t;           constexpr int kVmax = 4;
#elif TTVB_DTYPE == 5
using elem_t = unsigned long long; constexpr int kVmax = 2;
#else
#error "unknown TTVB_DTYPE"
#endif

#define TTVB_CAT2(a, b) a##b
#define TTVB_CAT(a, b)  TTVB_CAT2(a, b)

template<class Kernel>
static cudaError_t launch_tile(Kernel kern, const TileParams& P, const Launch& l, cudaStream_t stream)
{
  if (l.smem_bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem_bytes);
    if (e != cudaSuccess) return e;
  }
  return launch_k(kern, (unsigned)l.ctas, l.threads, l.smem_bytes, stream, P);
}

// (nu, ku) pairs that are instantiated: nu*ku = 8 loads in flight with 16-byte vectors, 16 with narrower ones
#define TTVB_BATCH_CASE(KERNEL, NU, KU) case (NU) * 100 + (KU): return launch_tile(KERNEL<T, V, NU, KU>, P, l, stream);

template<class T, int V>
static cudaError_t dispatch_batch(const TileParams& P, const Launch& l, cudaStream_t stream)
{
  const int key = l.nu * 100 + l.ku;
  constexpr bool wide = sizeof(T) * V >= 16;
  if (l.kernel == TTV_B200_KERNEL_DOT && l.peel) {
    if constexpr (V > 1) {
      switch (key) {
        TTVB_BATCH_CASE(ttv_dot_peel_kernel, 1, 8) TTVB_BATCH_CASE(ttv_dot_peel_kernel, 2, 4) TTVB_BATCH_CASE(ttv_dot_peel_kernel, 4, 2)
        default: return cudaErrorInvalidValue;
      }
    }
    return cudaErrorInvalidValue;
  }
  if (l.kernel == TTV_B200_KERNEL_COLX && l.warp == 2) {      // COLR: inner % V is a template parameter
    if constexpr (V > 1 && wide) {
      const int im = (int)(P.inner % (uint64_t)V);
#define TTVB_COLR_CASE(IM, NU, KU) case (IM) * 10000 + (NU) * 100 + (KU): return launch_tile(ttv_colr_kernel<T, V, IM, NU, KU>, P, l, stream);
      switch (im * 10000 + key) {
        TTVB_COLR_CASE(0, 1, 8) TTVB_COLR_CASE(0, 2, 4) TTVB_COLR_CASE(1, 1, 8) TTVB_COLR_CASE(1, 2, 4)
        default: break;
      }
      if constexpr (V == 2) {
        switch (im * 10000 + key) {
          TTVB_COLR_CASE(0, 4, 2) TTVB_COLR_CASE(1, 4, 2)
          default: break;
        }
      }
      if constexpr (V == 4) {
        switch (im * 10000 + key) {
          TTVB_COLR_CASE(2, 1, 8) TTVB_COLR_CASE(2, 2, 4) TTVB_COLR_CASE(3, 1, 8) TTVB_COLR_CASE(3, 2, 4)
          default: break;
        }
      }
#undef TTVB_COLR_CASE
    }
    return cudaErrorInvalidValue;
  }
  if (l.kernel == TTV_B200_KERNEL_COLX && l.warp) {
    if constexpr (V == 2 && sizeof(T) == 4) {      // 4-byte elements as 8-byte vectors: 2 phases, 4 accumulators per unit
      switch (key) {
        TTVB_BATCH_CASE(ttv_colw_kernel, 4, 4) TTVB_BATCH_CASE(ttv_colw_kernel, 2, 8) TTVB_BATCH_CASE(ttv_colw_kernel, 8, 2)
        default: return cudaErrorInvalidValue;
      }
    }
    if constexpr (V > 1 && wide) {
      switch (key) {
        TTVB_BATCH_CASE(ttv_colw_kernel, 1, 8) TTVB_BATCH_CASE(ttv_colw_kernel, 2, 4)
        default: break;
      }
      if constexpr (V == 2) {
        switch (key) {
          TTVB_BATCH_CASE(ttv_colw_kernel, 4, 2)
          default: break;
        }
      }
    }
    return cudaErrorInvalidValue;
  }
  if (l.kernel == TTV_B200_KERNEL_COLX) {
    if constexpr (V > 1 && wide) {
      switch (key) {
        TTVB_BATCH_CASE(ttv_colx_kernel, 1, 8) TTVB_BATCH_CASE(ttv_colx_kernel, 2, 4) TTVB_BATCH_CASE(ttv_colx_kernel, 4, 2)
        default: return cudaErrorInvalidValue;
      }
    }
    return cudaErrorInvalidValue;
  }
  if (l.bdirect) {      // b straight from L2 inside the batches: one batch shape, (1, 8)
    if (key != 108) return cudaErrorInvalidValue;
    if (l.kernel == TTV_B200_KERNEL_DOT) return launch_tile(ttv_dot_kernel<T, V, 1, 8, true>, P, l, stream);
    if (l.kernel == TTV_B200_KERNEL_COL) return launch_tile(ttv_col_kernel<T, V, 1, 8, true>, P, l, stream);
    return cudaErrorInvalidValue;
  }
  if (l.kernel == TTV_B200_KERNEL_DOT) {
    switch (key) {
      TTVB_BATCH_CASE(ttv_dot_kernel, 8, 1)
      default: break;
    }
    if constexpr (wide) {
      switch (key) {
        TTVB_BATCH_CASE(ttv_dot_kernel, 1, 8) TTVB_BATCH_CASE(ttv_dot_kernel, 2, 4) TTVB_BATCH_CASE(ttv_dot_kernel, 4, 2)
        TTVB_BATCH_CASE(ttv_dot_kernel, 2, 8) TTVB_BATCH_CASE(ttv_dot_kernel, 4, 4) TTVB_BATCH_CASE(ttv_dot_kernel, 8, 2)
        default: return cudaErrorInvalidValue;
      }
    } else {
      switch (key) {
        TTVB_BATCH_CASE(ttv_dot_kernel, 1, 16) TTVB_BATCH_CASE(ttv_dot_kernel, 2, 8) TTVB_BATCH_CASE(ttv_dot_kernel, 4, 4)
        TTVB_BATCH_CASE(ttv_dot_kernel, 8, 2)
        default: return cudaErrorInvalidValue;
      }
    }
  }
  if constexpr (wide) {
    switch (key) {
      TTVB_BATCH_CASE(ttv_col_kernel, 1, 8) TTVB_BATCH_CASE(ttv_col_kernel, 2, 4) TTVB_BATCH_CASE(ttv_col_kernel, 4, 2)
      TTVB_BATCH_CASE(ttv_col_kernel, 1, 16) TTVB_BATCH_CASE(ttv_col_kernel, 2, 8)
      default: return cudaErrorInvalidValue;
    }
  } else {
    switch (key) {
      TTVB_BATCH_CASE(ttv_col_kernel, 1, 16) TTVB_BATCH_CASE(ttv_col_kernel, 2, 8) TTVB_BATCH_CASE(ttv_col_kernel, 4, 4)
      TTVB_BATCH_CASE(ttv_col_kernel, 8, 2)
      default: return cudaErrorInvalidValue;
    }
  }
}

cudaError_t TTVB_CAT(tile_dtype_, TTVB_DTYPE)(const TileParams& P, const Launch& l, cudaStream_t stream)
{
  if constexpr (kVmax >= 4) if (l.vec == 4) return dispatch_batch<elem_t, 4>(P, l, stream);
  if constexpr (kVmax >= 2) if (l.vec == 2) return dispatch_batch<elem_t, 2>(P, l, stream);
  if (l.vec == 1) return dispatch_batch<elem_t, 1>(P, l, stream);
  return cudaErrorInvalidValue;
}

cudaError_t TTVB_CAT(reduce_dtype_, TTVB_DTYPE)(const void* ws, void* c, uint64_t n, uint32_t ksplit, bool accumulate,
                                                 uint64_t stride, int sm_count, cudaStream_t stream)
{
  if (ksplit >= 64 && n <= 8192) {       // few outputs, many partitions: a CTA per output (kernels.cuh)
    const uint64_t wide_blocks = std::min<uint64_t>(n, (uint64_t)sm_count * 8);
    return launch_k(ttv_reduce_wide_kernel<elem_t>, (unsigned)wide_blocks, 256u, 0, stream, static_cast<const elem_t*>(ws),
                    static_cast<elem_t*>(c), n, ksplit, accumulate ? 1u : 0u, stride);
  }
  const uint64_t blocks = std::min<uint64_t>((n + 255) / 256, (uint64_t)sm_count * 32);
  return launch_k(ttv_reduce_kernel<elem_t>, (unsigned)blocks, 256u, 0, stream, static_cast<const elem_t*>(ws), static_cast<elem_t*>(c), n,
                  ksplit, accumulate ? 1u : 0u, stride);
}

cudaError_t TTVB_CAT(stream_dtype_, TTVB_DTYPE)(const StreamParams& S, const Launch& l, cudaStream_t stream)
{
  // fibers (inner == 1) of even length are walked skewed: bank conflicts otherwise (stream_fibers_skewed)
  const bool skew = S.inner == 1 && (S.nq & 1) == 0;
  auto kern = skew ? ttv_stream_kernel<elem_t, 3, 4, true> : ttv_stream_kernel<elem_t, 3, 4, false>;
  if (l.threads > 256) {                     // a big slab alone in its stage: one CTA of up to 1024 threads per SM
    if (skew || l.threads > 1024) return cudaErrorInvalidValue;
    kern = l.stages == 3 ? ttv_stream_kernel<elem_t, 3, 1, false, 1024> : l.stages == 4 ? ttv_stream_kernel<elem_t, 4, 1, false, 1024>
         : ttv_stream_kernel<elem_t, 5, 1, false, 1024>;
    if (l.stages < 3 || l.stages > 5) return cudaErrorInvalidValue;
  } else if (l.stages == 4) kern = skew ? ttv_stream_kernel<elem_t, 4, 4, true> : ttv_stream_kernel<elem_t, 4, 4, false>;
  else if (l.stages == 5) kern = skew ? ttv_stream_kernel<elem_t, 5, 4, true> : ttv_stream_kernel<elem_t, 5, 4, false>;
  else if (l.stages != 3) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem_bytes);
  if (e != cudaSuccess) return e;
  return launch_k(kern, (unsigned)l.ctas, l.threads, l.smem_bytes, stream, S);
}

cudaError_t TTVB_CAT(dotf_dtype_, TTVB_DTYPE)(const DotfParams& D, const Launch& l, cudaStream_t stream)
{
  // b in registers (2 CTAs per SM) instead of shared memory (3 CTAs): measured on every DOTF shape of the named set
  // (profiles/r02_dotf_breg.txt): +5 % on fibers of 25 complex<double> (6 292 -> 6 618 GB/s, the slowest DOTF shape),
  // +2 % on 48, -1..-3 % on 24 / 40 and on the 8-byte types, +-0.5 % on 4-byte elements.  Rule: 16-byte elements, odd length.
  const char* br = std::getenv("TTV_B200_DOTF_BREG");
  const bool breg = (br && *br) ? std::atoi(br) != 0 : (sizeof(elem_t) == 16 && (D.nv & 1u));
  auto kern = breg ? ttv_dotf_kernel<elem_t, kVmax, true> : ttv_dotf_kernel<elem_t, kVmax, false>;
  if (l.smem_bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem_bytes);
    if (e != cudaSuccess) return e;
  }
  return launch_k(kern, (unsigned)l.ctas, l.threads, l.smem_bytes, stream, D);
}

template<int V>
static cudaError_t launch_scatter_vec(const ScatterParams& S, uint64_t ctas, uint64_t smem, cudaStream_t stream)
{
  if constexpr (V <= kVmax) {
    return launch_k(ttv_col_scatter_kernel<elem_t, V>, (unsigned)ctas, 256u, smem, stream, S);
  } else {
    return cudaErrorInvalidValue;
  }
}

cudaError_t TTVB_CAT(scatter_dtype_, TTVB_DTYPE)(const ScatterParams& S, int vec, uint64_t ctas, uint64_t smem, cudaStream_t stream)
{
  if (vec == 4) return launch_scatter_vec<4>(S, ctas, smem, stream);
  if (vec == 2) return launch_scatter_vec<2>(S, ctas, smem, stream);
  if (vec == 1) return launch_scatter_vec<1>(S, ctas, smem, stream);
  return cudaErrorInvalidValue;
}

template<int V>
static cudaError_t launch_exchange_vec(const ExchangeParams& E, uint32_t max_ctas, int sm_count, uint64_t smem, cudaStream_t stream)
{
  if constexpr (V <= kVmax) {
    auto kern = ttv_col_exchange_kernel<elem_t, V>;
    // the grid of the plain kernels (CTAs stride over the tiles); only the last `reducers` CTAs to arrive wait for the other
    // GPUs, one per SM unless capped -- far fewer than the device holds at once, so the rest always finds room
    ExchangeParams E2 = E;
    E2.reducers = max_ctas ? max_ctas : (uint32_t)sm_count;
    // Every CTA ends with a system-scope fence that waits for its peer stores to be acknowledged over NVLink (a few
    // microseconds): with one tile per CTA (the 64-per-SM grid of the plain kernels) that is 3-5 % of a CTA's life
    // (measured on 8 GPUs: 1.245 ms against 1.188 for the three-launch form).  Twelve CTAs per SM -- four rounds of the
    // three resident ones, several tiles each -- pay the fence once per several tiles and still end as evenly.
    const char* gm = std::getenv("TTV_B200_EXCHANGE_GRID_MULT");
    const uint64_t mult = (gm && *gm) ? (uint64_t)std::max(1, std::atoi(gm)) : 12;
    const uint64_t ctas = std::max<uint64_t>(1, std::min<uint64_t>(std::max<uint64_t>(E.S.tiles, 1), (uint64_t)sm_count * mult));
    return launch_k(kern, (unsigned)ctas, 256u, smem, stream, E2);
  } else {
    return cudaErrorInvalidValue;
  }
}

cudaError_t TTVB_CAT(exchange_dtype_, TTVB_DTYPE)(const ExchangeParams& E, int vec, uint32_t max_ctas, int sm_count, uint64_t smem, cudaStream_t stream)
{
  if (vec == 4) return launch_exchange_vec<4>(E, max_ctas, sm_count, smem, stream);
  if (vec == 2) return launch_exchange_vec<2>(E, max_ctas, sm_count, smem, stream);
  if (vec == 1) return launch_exchange_vec<1>(E, max_ctas, sm_count, smem, stream);
  return cudaErrorInvalidValue;
}

// TTV_B200_STRIDED_SCALAR: 1 = keep the thread-per-output form, 2 = take the vector form whenever the strides allow it
// (A/B measurements, tests of both forms); unset / 0 = the measured rule in strided_dtype_*
static int strided_form_forced()
{
  const char* e = std::getenv("TTV_B200_STRIDED_SCALAR");
  return (e && *e) ? std::atoi(e) : 0;
}

template<int V>
static cudaError_t launch_strided_vec(StridedParams S, int sm_count, cudaStream_t stream)
{
  if constexpr (V <= kVmax) {                           // (wider vectors than 16 bytes are never instantiated)
    S.n0p = (S.n[0] / (uint64_t)V + 31) / 32 * 32;      // whole warps per row
    const uint64_t lanes = (S.total / S.n[0]) * S.n0p;
    const uint64_t blocks = std::max<uint64_t>(1, std::min<uint64_t>((lanes + 255) / 256, (uint64_t)sm_count * 32));
    return launch_k(ttv_strided_vec_kernel<elem_t, V>, (unsigned)blocks, 256u, 0, stream, S);
  } else {
    return cudaErrorInvalidValue;
  }
}

template<int V>
static cudaError_t launch_strided_dot(const StridedParams& S, int sm_count, cudaStream_t stream)
{
  // lanes per fiber: one batch of loads each if a warp suffices, at most a warp
  const uint64_t kv = S.nq / (uint64_t)V;
  uint32_t G = 1;
  while (G < 32 && (uint64_t)G * strided_dot_ku<elem_t, V>() < kv) G *= 2;
  const uint64_t fibers_per_cta = 256 / G;
  const uint64_t blocks = std::max<uint64_t>(1, std::min<uint64_t>((S.total + fibers_per_cta - 1) / fibers_per_cta, (uint64_t)sm_count * 32));
  return launch_k(ttv_strided_dot_kernel<elem_t, V>, (unsigned)blocks, 256u, 0, stream, S, G);
}

cudaError_t TTVB_CAT(strided_dtype_, TTVB_DTYPE)(const StridedParams& S, int sm_count, cudaStream_t stream)
{
  if (S.wq == 1 && S.nq >= 8) {
    // q is the contiguous mode: lane groups along the fibers, with the widest vector the strides and addresses allow
    uint64_t align = (uint64_t)(reinterpret_cast<uintptr_t>(S.a) | reinterpret_cast<uintptr_t>(S.b)) / sizeof(elem_t);
    if ((reinterpret_cast<uintptr_t>(S.a) | reinterpret_cast<uintptr_t>(S.b)) % sizeof(elem_t)) align = 1;
    align |= S.nq;
    for (uint32_t d = 0; d < S.nfree; ++d) align |= S.wa[d];
    if constexpr (kVmax >= 4) if (align % 4 == 0) return launch_strided_dot<4>(S, sm_count, stream);
    if constexpr (kVmax >= 2) if (align % 2 == 0) return launch_strided_dot<2>(S, sm_count, stream);
    return launch_strided_dot<1>(S, sm_count, stream);
  }
  const int forced = strided_form_forced();
  if (kVmax > 1 && S.nfree >= 1 && S.wa[0] == 1 && S.wc[0] == 1 && forced != 1) {
    // the fastest free mode is contiguous in both tensors: V consecutive outputs per thread when everything else lines up
    const uintptr_t addr = reinterpret_cast<uintptr_t>(S.a) | reinterpret_cast<uintptr_t>(S.c);
    uint64_t align = (addr % sizeof(elem_t)) ? 1 : (uint64_t)(addr / sizeof(elem_t));
    align |= S.n[0] | S.wq;
    for (uint32_t d = 1; d < S.nfree; ++d) align |= S.wa[d] | S.wc[d];
    // Measured (profiles/r01_padded_strides.txt, session 4): 16-byte vectors of 4-byte elements gain 7-12 % over the
    // thread-per-output form (slices of a 256^4 fp32 tensor: 6.2-6.4 -> 6.8-6.9 TB/s) as long as rows come in whole warps
    // (see strided_kernel.cuh); 8-byte elements gain nothing either way -- their thread-per-output loop is not
    // issue-bound.  So: 4-byte elements, at most 1/8 of the lanes idle, enough rows to fill the machine.
    const uint64_t n0v = S.n[0] / 4, n0p = (n0v + 31) / 32 * 32;
    const bool pays = kVmax == 4 && (n0p - n0v) * 8 <= n0p && S.total / 4 >= (uint64_t)sm_count * 2048;
    if (align % kVmax == 0 && (pays || forced == 2)) return launch_strided_vec<kVmax>(S, sm_count, stream);
    if (forced == 2 && kVmax >= 4 && align % 2 == 0) return launch_strided_vec<2>(S, sm_count, stream);
  }
  const uint64_t blocks = std::max<uint64_t>(1, std::min<uint64_t>((S.total + 255) / 256, (uint64_t)sm_count * 32));
  return launch_k(ttv_strided_kernel<elem_t>, (unsigned)blocks, 256u, 0, stream, S);
}

cudaError_t TTVB_CAT(streamk_dtype_, TTVB_DTYPE)(const StreamkParams& K, const Launch& l, cudaStream_t stream)
{
  constexpr int kImaxAll = 64 / (int)sizeof(elem_t);                 // the widest row STREAMK takes: 64 bytes
  auto kern = ttv_streamk_kernel<elem_t, 3, kImaxAll>;
  if constexpr (kImaxAll >= 8) if (K.inner <= 4) kern = ttv_streamk_kernel<elem_t, 3, 4>;
  if constexpr (kImaxAll >= 16) if (K.inner > 4 && K.inner <= 8) kern = ttv_streamk_kernel<elem_t, 3, 8>;
  if ((int)K.inner > kImaxAll) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem_bytes);
  if (e != cudaSuccess) return e;
  return launch_k(kern, (unsigned)l.ctas, 256u, l.smem_bytes, stream, K);
}

cudaError_t TTVB_CAT(colf_tiny_dtype_, TTVB_DTYPE)(const ColfTinyParams& Y, const Launch& l, cudaStream_t stream)
{
  if constexpr (sizeof(elem_t) == 4) {
    return launch_k(ttv_colf_tiny_kernel<elem_t>, (unsigned)l.ctas, 256u, 0, stream, Y);
  } else {
    (void)Y; (void)l; (void)stream;
    return cudaErrorInvalidValue;
  }
}

cudaError_t TTVB_CAT(colf_dtype_, TTVB_DTYPE)(const ColfParams& F, const Launch& l, cudaStream_t stream)
{
  // a slab is at most one batch: b in registers, one pointer increment per item (TTV_B200_COLF_SHORT=0 keeps the general kernels)
  if constexpr (sizeof(elem_t) == 4) {
    if (l.short1 && l.pair) {
      if (l.stream) return launch_k(ttv_colfs_kernel<elem_t, 8, true, true>, (unsigned)l.ctas, 256u, 0, stream, F);
      return launch_k(ttv_colfs_kernel<elem_t, 8, false, true>, (unsigned)l.ctas, 256u, 0, stream, F);
    }
  }
  if constexpr (sizeof(elem_t) <= 8) {
    if (l.short1) {
      if (l.stream) return launch_k(ttv_colfs_kernel<elem_t, 8, true, false>, (unsigned)l.ctas, 256u, 0, stream, F);
      return launch_k(ttv_colfs_kernel<elem_t, 8, false, false>, (unsigned)l.ctas, 256u, 0, stream, F);
    }
  }
  if constexpr (sizeof(elem_t) == 4) {
    // rows of two elements: the form without selects, strip and second pass (TTV_B200_COLF_PAIR=0 keeps the general kernel)
    if (l.pair) {
      if (l.stream) return launch_k(ttv_colf2_kernel<elem_t, 8, true>, (unsigned)l.ctas, 256u, 0, stream, F);
      return launch_k(ttv_colf2_kernel<elem_t, 8, false>, (unsigned)l.ctas, 256u, 0, stream, F);
    }
  }
  if constexpr (sizeof(elem_t) <= 8) {
    if (l.stream) return launch_k(ttv_colf_kernel<elem_t, 8, true>, (unsigned)l.ctas, 256u, 0, stream, F);
    return launch_k(ttv_colf_kernel<elem_t, 8, false>, (unsigned)l.ctas, 256u, 0, stream, F);
  } else {
    (void)F; (void)l; (void)stream;
    return cudaErrorInvalidValue;                                    // 16-byte elements: every row is whole vectors
  }
}

cudaError_t TTVB_CAT(dotp_dtype_, TTVB_DTYPE)(const DotpParams& D, const Launch& l, cudaStream_t stream)
{
  if constexpr (sizeof(elem_t) <= 8) {
    if (l.ku == 4) return launch_k(ttv_dotp_kernel<elem_t, 4>, (unsigned)l.ctas, 256u, 0, stream, D);
    return launch_k(ttv_dotp_kernel<elem_t, 8>, (unsigned)l.ctas, 256u, 0, stream, D);
  } else {
    (void)D; (void)l; (void)stream;
    return cudaErrorInvalidValue;                                    // a fiber of two 16-byte elements is two vectors: DOTF
  }
}

cudaError_t TTVB_CAT(colt_dtype_, TTVB_DTYPE)(const CUtensorMap& map, const ColtParams& P, const Launch& l, cudaStream_t stream)
{
  constexpr int V = 16 / (int)sizeof(elem_t);
  auto kern = ttv_colt_kernel<elem_t, V>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem_bytes);
  if (e != cudaSuccess) return e;
  return launch_k(kern, (unsigned)l.ctas, (unsigned)kColtThreads, l.smem_bytes, stream, map, P);
}

cudaError_t TTVB_CAT(fill_dtype_, TTVB_DTYPE)(void* x, uint64_t first, uint64_t count, uint64_t seed, int sm_count, cudaStream_t stream)
{
  const unsigned blocks = (unsigned)std::min<uint64_t>((count + 255) / 256, (uint64_t)sm_count * 32);
  return launch_k(ttv_fill_kernel<elem_t>, blocks, 256u, 0, stream, static_cast<elem_t*>(x), first, count, seed);
}
#endif

} // namespace ttvb
