// Microbenchmark: MUFU.EX2 (ex2.approx.ftz.f32) throughput per SM as a function of resident warps, alone and with
// packed-FMA work interleaved (what the attention softmax does around its exponentials).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I panacea_b200/csrc -o tools/ubench/mufu_rate tools/ubench/mufu_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace pn;

template <int FMA_PER_EX2>
__global__ void __launch_bounds__(1024, 1) k(int iters, long long* out, float* sink, float seed) {
  float x[8], y[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { x[j] = seed + threadIdx.x * 1e-3f + j; y[j] = 0.f; }
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      x[j] = ex2_approx(x[j]) * 0.5f - 1.0f;          // dependent per chain, 8 independent chains (the FMA keeps x bounded)
#pragma unroll
      for (int f = 0; f < FMA_PER_EX2; ++f) y[j] = fmaf(y[j], 0.999f, x[j]);
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) acc += x[j] + y[j];
  if (acc == 12345.678f) sink[0] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int F>
static void run(int warps) {
  long long* d; float* s;
  cudaMalloc(&d, 16); cudaMalloc(&s, 16);
  const int iters = 20000;
  k<F><<<148, warps * 32>>>(iters, d, s, 0.1f);
  k<F><<<148, warps * 32>>>(iters, d, s, 0.1f);
  cudaError_t e = cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%d FMA per EX2, %2d warps: %6.2f EX2/clk/SM  (%s)\n", F, warps, (double)iters * 8 * 32 * warps / (double)h, cudaGetErrorString(e));
  cudaFree(d); cudaFree(s);
}

int main() {
  for (int w : {4, 8, 16, 32}) run<0>(w);
  for (int w : {8, 16}) run<2>(w);
  for (int w : {8, 16}) run<4>(w);
  return 0;
}
