// Micro-benchmark (developer tool): do warp shuffles share the shared-memory data pipe?  Three kernels with 16 resident
// warps per SM (2 blocks x 8): A = 64-bit conflict-free shared stores + loads, B = shuffles, C = both interleaved.
// If time(C) ~ max(A, B) the two are separate resources; if ~ A + B they share one.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o shfl_vs_smem shfl_vs_smem.cu && ./shfl_vs_smem
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256, 2) k(float2* out, int iters) {
  __shared__ float2 buf[8][512];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = make_float2(lane + i, lane - i);
  float2* b = buf[warp];
  for (int it = 0; it < iters; ++it) {
    if (MODE & 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) b[lane + 32 * i] = v[i];
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 t = b[((lane + 1) & 31) + 32 * i];
        v[i].x += t.x;
        v[i].y += t.y;
      }
      __syncwarp();
    }
    if (MODE & 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i].x += __shfl_sync(0xffffffffu, v[i].y, (lane + 3) & 31);
        v[i].y += __shfl_sync(0xffffffffu, v[i].x, (lane + 5) & 31);
      }
    }
  }
  float2 s = make_float2(0, 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) { s.x += v[i].x; s.y += v[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
float run(float2* out, int iters) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<296, 256>>>(out, 10);
  cudaEventRecord(a);
  k<MODE><<<296, 256>>>(out, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  float2* out; cudaMalloc(&out, 296 * 256 * sizeof(float2));
  const int iters = 20000;
  const float a = run<1>(out, iters), b = run<2>(out, iters), c = run<3>(out, iters);
  // per iteration and SM: A = 16 warps x (8 STS.64 + 8 LDS.64) = 16 x 32 wavefronts; B = 16 warps x 16 SHFL
  printf("A smem  : %.3f ms  (%.2f ns per warp-iteration-SM-slot)\n", a, a * 1e6 / iters);
  printf("B shfl  : %.3f ms\n", b);
  printf("C both  : %.3f ms   A+B = %.3f, max = %.3f\n", c, a + b, a > b ? a : b);
  printf("clk/SM/iter at 1.9 GHz: A %.0f (512 wavefronts), B %.0f (256 shuffles), C %.0f\n", a * 1.9e6 / iters, b * 1.9e6 / iters, c * 1.9e6 / iters);
  return cudaGetLastError() != cudaSuccess;
}
