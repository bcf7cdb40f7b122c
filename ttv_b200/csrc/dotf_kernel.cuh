// dotf_kernel.cuh -- DOTF: mode q contiguous (inner == 1) with SHORT fibers, read as one flat stream.
//
// With fibers of a few 16-byte vectors (n_q = 4 ... 160 floats), giving every fiber its own lane group makes a warp-wide
// load touch many short separate pieces and leaves lanes idle whenever the vectors of a fiber do not divide evenly among
// its lanes (84 floats = 21 vectors on 4 lanes).  But the fibers of the view A[outer][n_q] lie back to back: the
// whole of A is ONE contiguous stream.  So a warp takes a chunk of F whole fibers (F * nv <= 256 vectors, nv = n_q / V),
// its lanes load CONSECUTIVE vectors (512 contiguous bytes per instruction, 8 instructions in flight, every slot
// useful), each vector is reduced with the matching vector of b to one partial product sum, the 256 partials go to the
// warp's own 256-entry strip of shared memory, and lane f adds up the nv partials of fiber f.  Warps never wait for each
// other (only __syncwarp), and the loads of the next chunk are already in flight while the partials of the current one
// are summed.  Because a chunk starts on a fiber boundary, the position of a lane's vector inside its fiber -- and so
// the vector of b it needs -- is the same in every chunk.
//
// Replaces the same reference code as ttv_dot_kernel (gemv_row / dot, detail/matrix_times_vector.h:51-91,264-295, inside
// the loop nest of detail/tensor_times_vector.h:189-324).  Requires n_q % V == 0 and a 16-byte aligned A and b.
#pragma once

#include "kernels.cuh"

namespace ttvb {

struct DotfParams {
  const void* a;
  const void* b;
  void*       c;
  uint64_t outer;           // fibers
  uint64_t chunks;          // ceil(outer / fw)
  uint32_t nq;
  uint32_t nv;              // vectors per fiber
  uint32_t fw;              // fibers per chunk: fw * nv <= 256
  uint32_t accumulate;
  uint32_t lpf;             // lanes per fiber in the summation (power of two, lpf * min(fw, 32) <= 32)
};

// BREG: the KU vectors of b a lane needs are the same in every chunk (see above), so they can live in registers instead of
// being fetched from shared memory for every chunk -- 32 more registers (two CTAs per SM instead of three) against a fifth
// fewer shared-memory wavefronts in a kernel whose load/store pipe is the busiest unit (ncu: l1tex 93 % on complex<double>).
template<class T, int V, bool BREG = false>
__global__ void __launch_bounds__(256, BREG ? 2 : 3)
ttv_dotf_kernel(const DotfParams P)
{
  pdl_prologue();
  constexpr int KU = 8;                                  // vector loads in flight per lane; 32 * KU vectors per chunk
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sbv  = reinterpret_cast<T*>(smem_raw);              // [nq]: b
  T* part = sbv + P.nq;                                  // [8 warps][256]: partial sums, one per vector of the chunk

  const T* __restrict__ A = static_cast<const T*>(P.a);
  const T* __restrict__ B = static_cast<const T*>(P.b);
  T* __restrict__       C = static_cast<T*>(P.c);

  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t nv = P.nv, fw = P.fw, nq = P.nq;
  T* pw = part + (size_t)warp * (32 * KU);

  for (uint32_t j = threadIdx.x; j < nq; j += blockDim.x) sbv[j] = B[j];
  __syncthreads();

  // vector s*32 + lane of a chunk sits at position pos[s] of its fiber -- the same in every chunk
  uint32_t pos[KU];
  {
    uint32_t p = lane % nv;
    const uint32_t step = 32u % nv;
#pragma unroll
    for (int s = 0; s < KU; ++s) {
      pos[s] = p;
      p += step;
      if (p >= nv) p -= nv;
    }
  }

  Vec<T, V> breg[BREG ? KU : 1];
  if constexpr (BREG) {
#pragma unroll
    for (int s = 0; s < KU; ++s) breg[s] = *reinterpret_cast<const Vec<T, V>*>(sbv + pos[s] * V);
  }

  const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
  uint64_t chunk = (uint64_t)blockIdx.x * (blockDim.x >> 5) + warp;

  Vec<T, V> v[KU];
  auto load_chunk = [&](uint64_t ch) {
    const uint64_t f0 = ch * fw;
    const uint32_t nf = (uint32_t)min((uint64_t)fw, P.outer - f0);
    const uint32_t nvalid = nf * nv;
    const T* base = A + f0 * nq + (uint64_t)lane * V;
    if (nvalid == 32 * KU) {
#pragma unroll
      for (int s = 0; s < KU; ++s) v[s] = load_a<T, V>(base + (size_t)s * 32 * V, false);
    } else {
#pragma unroll
      for (int s = 0; s < KU; ++s)
        v[s] = (s * 32 + lane < nvalid) ? load_a<T, V>(base + (size_t)s * 32 * V, false) : zero_vec<T, V>();
    }
  };

  if (chunk < P.chunks) load_chunk(chunk);
  while (chunk < P.chunks) {
    // one partial per vector
#pragma unroll
    for (int s = 0; s < KU; ++s) {
      Vec<T, V> bv;
      if constexpr (BREG) bv = breg[s];
      else bv = *reinterpret_cast<const Vec<T, V>*>(sbv + pos[s] * V);
      T p = Num<T>::zero();
#pragma unroll
      for (int e = 0; e < V; ++e) p = Num<T>::madd(v[s].e[e], bv.e[e], p);
      pw[s * 32 + lane] = p;
    }
    __syncwarp();
    const uint64_t next = chunk + warps;
    if (next < P.chunks) load_chunk(next);               // in flight while this chunk's partials are summed

    const uint64_t f0 = chunk * fw;
    const uint32_t nf = (uint32_t)min((uint64_t)fw, P.outer - f0);
    if (P.lpf == 1) {
      // a lane per fiber
      for (uint32_t f = lane; f < nf; f += 32) {
        // even nv: start at a rotated position so that neighbouring lanes hit different banks
        uint32_t j = (nv & 1u) ? 0u : f % nv;
        const T* pf = pw + (size_t)f * nv;
        T sum = Num<T>::zero();
        for (uint32_t t = 0; t < nv; ++t) {
          sum = Num<T>::add(sum, pf[j]);
          j = (j + 1 == nv) ? 0u : j + 1;
        }
        T* out = C + f0 + f;
        *out = P.accumulate ? Num<T>::add(*out, sum) : sum;
      }
    } else {
      // few long fibers per chunk: lpf lanes share a fiber (contiguous pieces), butterfly inside the group
      const uint32_t lpf = P.lpf, seg = (nv + lpf - 1) / lpf;
      const uint32_t r = lane % lpf, fl = lane / lpf, fpr = 32 / lpf;
      for (uint32_t fb = 0; fb < fw; fb += fpr) {          // warp-uniform trip count
        const uint32_t f = fb + fl;
        T sum = Num<T>::zero();
        if (f < nf && r * seg < nv) {
          const T* pf = pw + (size_t)f * nv + r * seg;
          const uint32_t len = min(seg, nv - r * seg);
          uint32_t j = (seg & 1u) ? 0u : lane % len;       // even stride between lanes: rotate the start (banks)
          for (uint32_t t = 0; t < len; ++t) {
            sum = Num<T>::add(sum, pf[j]);
            j = (j + 1 == len) ? 0u : j + 1;
          }
        }
        for (uint32_t h = lpf >> 1; h > 0; h >>= 1) sum = Num<T>::add(sum, shfl_xor_elem(sum, (int)h));
        if (f < nf && r == 0) {
          T* out = C + f0 + f;
          *out = P.accumulate ? Num<T>::add(*out, sum) : sum;
        }
      }
    }
    __syncwarp();
    chunk = next;
  }
}

} // namespace ttvb
