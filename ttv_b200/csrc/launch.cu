// launch.cu -- instantiates the kernels of kernels.cuh per (element type, vector width, unroll) and launches them.
#include "launch.h"
#include "kernels.cuh"

#include <atomic>

namespace ttvb {

static std::atomic<uint64_t> g_launches{0};
uint64_t launch_count() { return g_launches.load(); }

template<class Kernel>
static cudaError_t launch_tile(Kernel kern, const TileParams& P, const Launch& l, cudaStream_t stream)
{
  if (l.smem_bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem_bytes);
    if (e != cudaSuccess) return e;
  }
  kern<<<(unsigned)l.ctas, l.threads, l.smem_bytes, stream>>>(P);
  g_launches.fetch_add(1);
  return cudaGetLastError();
}

template<class T, int V>
static cudaError_t dispatch_ku(const TileParams& P, const Launch& l, cudaStream_t stream)
{
  const bool dot = l.kernel == TTV_B200_KERNEL_DOT;
  switch (l.ku) {
    case 1: case 2: case 4:
      return dot ? launch_tile(ttv_dot_kernel<T, V, 4>, P, l, stream) : launch_tile(ttv_col_kernel<T, V, 4>, P, l, stream);
    default:
      return dot ? launch_tile(ttv_dot_kernel<T, V, 8>, P, l, stream) : launch_tile(ttv_col_kernel<T, V, 8>, P, l, stream);
  }
}

template<class T, int VMAX>
static cudaError_t dispatch_vec(const TileParams& P, const Launch& l, cudaStream_t stream)
{
  if constexpr (VMAX >= 4) if (l.vec == 4) return dispatch_ku<T, 4>(P, l, stream);
  if constexpr (VMAX >= 2) if (l.vec == 2) return dispatch_ku<T, 2>(P, l, stream);
  if (l.vec == 1) return dispatch_ku<T, 1>(P, l, stream);
  return cudaErrorInvalidValue;
}

template<class T>
static cudaError_t run_reduce(const void* ws, void* c, uint64_t n, uint32_t ksplit, bool accumulate, int sm_count, cudaStream_t stream)
{
  const uint64_t blocks = std::min<uint64_t>((n + 255) / 256, (uint64_t)sm_count * 32);
  ttv_reduce_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const T*>(ws), static_cast<T*>(c), n, ksplit, accumulate ? 1u : 0u);
  g_launches.fetch_add(1);
  return cudaGetLastError();
}

cudaError_t launch_view(int dtype, const View& v, const Launch& l, const void* a, const void* b, void* c,
                        void* workspace, bool accumulate, int sm_count, cudaStream_t stream)
{
  TileParams P;
  P.a = a; P.b = b;
  P.c = l.ksplit > 1 ? workspace : c;
  P.outer = v.outer; P.nq = v.nq; P.inner = v.inner;
  P.kchunk = l.kchunk;
  P.itiles = l.itiles; P.otiles = l.otiles; P.tiles = l.tiles;
  P.tx = l.tx; P.ty = l.ty; P.to = l.to;
  P.ksplit = l.ksplit; P.kb = l.kb;
  P.accumulate = accumulate ? 1u : 0u;

  cudaError_t e;
  switch (dtype) {
    case TTV_B200_F32:  e = dispatch_vec<float, 4>(P, l, stream); break;
    case TTV_B200_F64:  e = dispatch_vec<double, 2>(P, l, stream); break;
    case TTV_B200_C64:  e = dispatch_vec<cf32, 2>(P, l, stream); break;
    case TTV_B200_C128: e = dispatch_vec<cf64, 1>(P, l, stream); break;
    case TTV_B200_I32:  e = dispatch_vec<uint32_t, 4>(P, l, stream); break;
    case TTV_B200_I64:  e = dispatch_vec<unsigned long long, 2>(P, l, stream); break;
    default: return cudaErrorInvalidValue;
  }
  if (e != cudaSuccess || l.ksplit <= 1) return e;

  const uint64_t n = v.outer * v.inner;
  switch (dtype) {
    case TTV_B200_F32:  return run_reduce<float>(workspace, c, n, l.ksplit, accumulate, sm_count, stream);
    case TTV_B200_F64:  return run_reduce<double>(workspace, c, n, l.ksplit, accumulate, sm_count, stream);
    case TTV_B200_C64:  return run_reduce<cf32>(workspace, c, n, l.ksplit, accumulate, sm_count, stream);
    case TTV_B200_C128: return run_reduce<cf64>(workspace, c, n, l.ksplit, accumulate, sm_count, stream);
    case TTV_B200_I32:  return run_reduce<uint32_t>(workspace, c, n, l.ksplit, accumulate, sm_count, stream);
    case TTV_B200_I64:  return run_reduce<unsigned long long>(workspace, c, n, l.ksplit, accumulate, sm_count, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_fill(int dtype, void* x, uint64_t first, uint64_t count, uint64_t seed, int sm_count, cudaStream_t stream)
{
  if (count == 0) return cudaSuccess;
  const unsigned blocks = (unsigned)std::min<uint64_t>((count + 255) / 256, (uint64_t)sm_count * 32);
  switch (dtype) {
    case TTV_B200_F32:  ttv_fill_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<float*>(x), first, count, seed); break;
    case TTV_B200_F64:  ttv_fill_kernel<double><<<blocks, 256, 0, stream>>>(static_cast<double*>(x), first, count, seed); break;
    case TTV_B200_C64:  ttv_fill_kernel<cf32><<<blocks, 256, 0, stream>>>(static_cast<cf32*>(x), first, count, seed); break;
    case TTV_B200_C128: ttv_fill_kernel<cf64><<<blocks, 256, 0, stream>>>(static_cast<cf64*>(x), first, count, seed); break;
    case TTV_B200_I32:  ttv_fill_kernel<uint32_t><<<blocks, 256, 0, stream>>>(static_cast<uint32_t*>(x), first, count, seed); break;
    case TTV_B200_I64:  ttv_fill_kernel<unsigned long long><<<blocks, 256, 0, stream>>>(static_cast<unsigned long long*>(x), first, count, seed); break;
    default: return cudaErrorInvalidValue;
  }
  g_launches.fetch_add(1);
  return cudaGetLastError();
}

} // namespace ttvb
