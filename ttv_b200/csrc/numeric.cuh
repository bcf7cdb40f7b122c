// numeric.cuh -- element types, multiply-add and vector loads for the TTV kernels (sm_100a).
//
// The path is an HBM-bound GEMV: 2 FLOP per element read.  Tensor cores are deliberately not used; what matters
// here is 128-bit coalesced loads, enough of them in flight, and not touching any byte of A twice.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace ttvb {

// ---- element traits ------------------------------------------------------------------------------------------
// Device-side element types: float, double, float2 (complex<float>), double2 (complex<double>), int32, int64.
// Integer arithmetic is done on the unsigned twin so that overflow wraps (bit-exact with any summation order).
struct cf32 { float  re, im; };
struct cf64 { double re, im; };

template<class T> struct Num;

template<> struct Num<float> {
  static __device__ __forceinline__ float zero() { return 0.f; }
  static __device__ __forceinline__ float madd(float a, float b, float acc) { return fmaf(a, b, acc); }
  static __device__ __forceinline__ float add(float a, float b) { return a + b; }
};
template<> struct Num<double> {
  static __device__ __forceinline__ double zero() { return 0.0; }
  static __device__ __forceinline__ double madd(double a, double b, double acc) { return fma(a, b, acc); }
  static __device__ __forceinline__ double add(double a, double b) { return a + b; }
};
template<> struct Num<uint32_t> {
  static __device__ __forceinline__ uint32_t zero() { return 0u; }
  static __device__ __forceinline__ uint32_t madd(uint32_t a, uint32_t b, uint32_t acc) { return a * b + acc; }
  static __device__ __forceinline__ uint32_t add(uint32_t a, uint32_t b) { return a + b; }
};
template<> struct Num<unsigned long long> {
  static __device__ __forceinline__ unsigned long long zero() { return 0ull; }
  static __device__ __forceinline__ unsigned long long madd(unsigned long long a, unsigned long long b, unsigned long long acc) { return a * b + acc; }
  static __device__ __forceinline__ unsigned long long add(unsigned long long a, unsigned long long b) { return a + b; }
};
// complex: plain (non-conjugated) product, real and imaginary parts accumulated separately
// (reference matrix_times_vector.h:65,124 use operator* of std::complex).
template<> struct Num<cf32> {
  static __device__ __forceinline__ cf32 zero() { return cf32{0.f, 0.f}; }
  static __device__ __forceinline__ cf32 madd(cf32 a, cf32 b, cf32 acc) {
    acc.re = fmaf(a.re, b.re, acc.re); acc.re = fmaf(-a.im, b.im, acc.re);
    acc.im = fmaf(a.re, b.im, acc.im); acc.im = fmaf(a.im, b.re, acc.im);
    return acc;
  }
  static __device__ __forceinline__ cf32 add(cf32 a, cf32 b) { return cf32{a.re + b.re, a.im + b.im}; }
};
template<> struct Num<cf64> {
  static __device__ __forceinline__ cf64 zero() { return cf64{0.0, 0.0}; }
  static __device__ __forceinline__ cf64 madd(cf64 a, cf64 b, cf64 acc) {
    acc.re = fma(a.re, b.re, acc.re); acc.re = fma(-a.im, b.im, acc.re);
    acc.im = fma(a.re, b.im, acc.im); acc.im = fma(a.im, b.re, acc.im);
    return acc;
  }
  static __device__ __forceinline__ cf64 add(cf64 a, cf64 b) { return cf64{a.re + b.re, a.im + b.im}; }
};

// ---- vectors of V elements, loaded with one instruction ---------------------------------------------------------
template<class T, int V> struct alignas(sizeof(T) * V) Vec { T e[V]; };

// Streaming load of A: read-only path, do not allocate in L1 (every byte of A is used exactly once), so that
// L1/shared memory stays available for b and the reduction scratch.
template<int BYTES> struct Ld;
template<> struct Ld<16> {
  static __device__ __forceinline__ void nc(void* dst, const void* src) {
    uint32_t x, y, z, w;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "l"(src));
    uint32_t* d = reinterpret_cast<uint32_t*>(dst); d[0] = x; d[1] = y; d[2] = z; d[3] = w;
  }
};
template<> struct Ld<8> {
  static __device__ __forceinline__ void nc(void* dst, const void* src) {
    uint32_t x, y;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "l"(src));
    uint32_t* d = reinterpret_cast<uint32_t*>(dst); d[0] = x; d[1] = y;
  }
};
template<> struct Ld<4> {
  static __device__ __forceinline__ void nc(void* dst, const void* src) {
    uint32_t x;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(x) : "l"(src));
    *reinterpret_cast<uint32_t*>(dst) = x;
  }
};
// 32-byte vectors (V=2 of a 16-byte element) are two 16-byte loads
template<> struct Ld<32> {
  static __device__ __forceinline__ void nc(void* dst, const void* src) {
    Ld<16>::nc(dst, src);
    Ld<16>::nc(reinterpret_cast<char*>(dst) + 16, reinterpret_cast<const char*>(src) + 16);
  }
};

// A 16-byte load under a predicate, zeros otherwise: one predicated instruction, no branch around the asm statement (a
// batch whose loads are all conditional -- a slab shorter than the batch -- keeps them back to back).  NA: L1::no_allocate.
template<bool NA>
__device__ __forceinline__ void ld16_if(void* dst, const void* src, bool ok)
{
  uint32_t x, y, z, w;
  if constexpr (NA)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\tmov.b32 %2, 0;\n\tmov.b32 %3, 0;\n\t"
                 "@p ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];\n\t}"
                 : "=&r"(x), "=&r"(y), "=&r"(z), "=&r"(w) : "l"(src), "r"((uint32_t)ok));
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\tmov.b32 %2, 0;\n\tmov.b32 %3, 0;\n\t"
                 "@p ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];\n\t}"
                 : "=&r"(x), "=&r"(y), "=&r"(z), "=&r"(w) : "l"(src), "r"((uint32_t)ok));
  uint32_t* d = reinterpret_cast<uint32_t*>(dst); d[0] = x; d[1] = y; d[2] = z; d[3] = w;
}

template<class T, int V>
__device__ __forceinline__ Vec<T, V> load_stream(const T* p) {
  Vec<T, V> v;
  Ld<sizeof(T) * V>::nc(&v, p);
  return v;
}

// ---- programmatic dependent launch (sm_90+) -------------------------------------------------------------------------
// Every kernel of this library starts with pdl_prologue(): it lets the NEXT kernel of the stream start placing its CTAs as
// soon as all CTAs of this one have started (griddepcontrol.launch_dependents), and it waits until the PREVIOUS kernel of
// the stream has completed and flushed its memory before touching anything (griddepcontrol.wait) -- so the semantics of a
// stream are unchanged (a product may read what the product before it wrote, as in a ttvs chain), but the launch latency
// and the CTA ramp of a kernel overlap the tail of the one before it.  Both instructions are no-ops for kernels launched
// without the programmatic-serialization attribute (launch.cu, TTV_B200_PDL).
__device__ __forceinline__ void pdl_prologue()
{
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---- synthetic data: identical to oracle/ttv_oracle.c ttv_oracle_fill -------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ double unit_pm1(uint64_t u) {
  return (double)(u >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

} // namespace ttvb
