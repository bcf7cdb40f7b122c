// tools/probe/tail_probe.cu -- where do the ~10 us of fixed cost per launch go?  A column-GEMV-shaped streaming kernel
// (A[rows][cols] float, y[c] = sum_r A[r][c] * b[r]) whose CTAs record %globaltimer at start / after staging b / at end.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probe/tail_probe tools/probe/tail_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

template<int L, bool BFIRST>
__global__ void __launch_bounds__(256, (L > 8 ? 2 : 3))
colk(const float4* __restrict__ A, const float* __restrict__ B, float4* __restrict__ C, int rows, long cols4, long tiles,
     unsigned long long* ts, unsigned* counter)
{
  __shared__ float sb[4096];
  __shared__ long s_tile;
  unsigned long long t0 = gtime();
  if (BFIRST) { for (int j = threadIdx.x; j < rows; j += blockDim.x) sb[j] = B[j]; __syncthreads(); }
  unsigned long long t1 = gtime();
  long tile = blockIdx.x;
  while (tile < tiles) {
    const long c = tile * blockDim.x + threadIdx.x;
    float4 acc = make_float4(0, 0, 0, 0);
    if (c < cols4) {
      const float4* p = A + c;
      int r = 0;
      if (!BFIRST) {
        // first batch's loads go out before b is staged
        float4 v[L];
#pragma unroll
        for (int s = 0; s < L; ++s) v[s] = p[(long)(r + s) * cols4];
        for (int j = threadIdx.x; j < rows; j += blockDim.x) sb[j] = B[j];
        __syncthreads();
#pragma unroll
        for (int s = 0; s < L; ++s) { float bb = sb[r + s]; acc.x += v[s].x * bb; acc.y += v[s].y * bb; acc.z += v[s].z * bb; acc.w += v[s].w * bb; }
        r += L;
      }
      for (; r + L <= rows; r += L) {
        float4 v[L];
#pragma unroll
        for (int s = 0; s < L; ++s) v[s] = p[(long)(r + s) * cols4];
#pragma unroll
        for (int s = 0; s < L; ++s) { float bb = sb[r + s]; acc.x += v[s].x * bb; acc.y += v[s].y * bb; acc.z += v[s].z * bb; acc.w += v[s].w * bb; }
      }
      C[c] = acc;
    }
    if (counter) {      // dynamic tile scheduler
      __syncthreads();
      if (threadIdx.x == 0) s_tile = gridDim.x + atomicAdd(counter, 1u);
      __syncthreads();
      tile = s_tile;
    } else tile += gridDim.x;
  }
  unsigned long long t2 = gtime();
  if (threadIdx.x == 0) { ts[blockIdx.x * 3] = t0; ts[blockIdx.x * 3 + 1] = t1; ts[blockIdx.x * 3 + 2] = t2; }
}

__global__ void stamp(unsigned long long* t) { *t = gtime(); }

template<int L, bool BFIRST>
void run(const char* name, float4** As, int ncopy, const float* B, float4* C, int rows, long cols4, int threads, int grid, bool dyn)
{
  const long tiles = (cols4 + threads - 1) / threads;
  if (grid <= 0 || grid > tiles) grid = (int)tiles;
  unsigned long long* ts; cudaMalloc(&ts, sizeof(unsigned long long) * (3 * grid + 2));
  unsigned* counter; cudaMalloc(&counter, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9, sum = 0; int reps = 12;
  std::vector<unsigned long long> h(3 * grid + 2);
  double st_spread = 0, en_spread = 0, first_start = 0, b_stage = 0, last_to_after = 0, dur_g = 0;
  for (int i = 0; i < reps + 3; ++i) {
    cudaMemsetAsync(counter, 0, 4);
    stamp<<<1, 1>>>(ts + 3 * grid);
    cudaEventRecord(e0);
    colk<L, BFIRST><<<grid, threads>>>(As[i % ncopy], B, C, rows, cols4, tiles, ts, dyn ? counter : nullptr);
    cudaEventRecord(e1);
    stamp<<<1, 1>>>(ts + 3 * grid + 1);
    cudaEventSynchronize(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (i >= 3) {
      best = std::min(best, ms); sum += ms;
      cudaMemcpy(h.data(), ts, sizeof(unsigned long long) * (3 * grid + 2), cudaMemcpyDeviceToHost);
      unsigned long long s0 = ~0ull, s1 = 0, e_0 = ~0ull, e_1 = 0; double bs = 0;
      for (int c = 0; c < grid; ++c) { s0 = std::min(s0, h[3 * c]); s1 = std::max(s1, h[3 * c]); e_0 = std::min(e_0, h[3 * c + 2]); e_1 = std::max(e_1, h[3 * c + 2]); bs += (double)(h[3 * c + 1] - h[3 * c]); }
      st_spread += (double)(s1 - s0); en_spread += (double)(e_1 - e_0); first_start += (double)(s0 - h[3 * grid]); b_stage += bs / grid;
      last_to_after += (double)(h[3 * grid + 1] - e_1); dur_g += (double)(e_1 - s0);
    }
  }
  const double bytes = 16.0 * rows * cols4;
  printf("%-34s grid %5d x %3d  L=%2d  event med/best %.2f / %.2f us  %.0f GB/s | prev-stamp->first CTA %.2f us, start spread %.2f, b stage %.2f, "
         "first start->last end %.2f, end spread %.2f, last end->next-stamp %.2f\n",
         name, grid, threads, L, sum / reps * 1e3, best * 1e3, bytes / (sum / reps) / 1e6, first_start / reps / 1e3, st_spread / reps / 1e3,
         b_stage / reps / 1e3, dur_g / reps / 1e3, en_spread / reps / 1e3, last_to_after / reps / 1e3);
  cudaFree(ts); cudaFree(counter);
}

int main(int argc, char** argv)
{
  const int rows = argc > 1 ? atoi(argv[1]) : 512;
  const long cols4 = 65536;
  const int ncopy = 4;
  float4* As[ncopy];
  for (int i = 0; i < ncopy; ++i) { cudaMalloc(&As[i], 16ull * rows * cols4); cudaMemset(As[i], 0, 16ull * rows * cols4); }
  float* B; cudaMalloc(&B, 4 * 4096); cudaMemset(B, 0, 4 * 4096);
  float4* C; cudaMalloc(&C, 16 * cols4);
  printf("rows %d, %ld MB per launch\n", rows, (long)(16ull * rows * cols4 >> 20));
  run<8, true>("L8 b-first 256 CTAs", As, ncopy, B, C, rows, cols4, 256, 0, false);
  run<16, true>("L16 b-first 256 CTAs", As, ncopy, B, C, rows, cols4, 256, 0, false);
  run<16, false>("L16 A-first 256 CTAs", As, ncopy, B, C, rows, cols4, 256, 0, false);
  run<8, false>("L8 A-first 256 CTAs", As, ncopy, B, C, rows, cols4, 256, 0, false);
  run<16, true>("L16 b-first 512 x 128thr", As, ncopy, B, C, rows, cols4, 128, 0, false);
  run<16, false>("L16 A-first 512 x 128thr", As, ncopy, B, C, rows, cols4, 128, 0, false);
  run<16, false>("L16 A-first 1024 x 64thr", As, ncopy, B, C, rows, cols4, 64, 0, false);
  run<8, false>("L8 A-first 1024 x 64thr", As, ncopy, B, C, rows, cols4, 64, 0, false);
  run<16, false>("L16 A-first 2048 x 32thr", As, ncopy, B, C, rows, cols4, 32, 0, false);
  run<16, false>("L16 A-first 64thr grid 592 dyn", As, ncopy, B, C, rows, cols4, 64, 592, true);
  run<16, false>("L16 A-first 32thr grid 1184 dyn", As, ncopy, B, C, rows, cols4, 32, 1184, true);
  run<16, true>("L16 b-first 96thr (683)", As, ncopy, B, C, rows, cols4, 96, 0, false);
  run<16, true>("L16 b-first 224thr (293)", As, ncopy, B, C, rows, cols4, 224, 0, false);
  return 0;
}
