// Microbenchmark: tcgen05.ld / tcgen05.st throughput of one SM (all 148 at once) as a function of the number of reading
// warps: every warp reads (writes) its lane quarter, 32 columns (4 KB) per instruction, in a loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I panacea_b200/csrc -o tools/ubench/tmem_rate tools/ubench/tmem_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace pn;

template <int MODE>      // 0 = ld x32, 1 = st x32, 2 = ld x16
__global__ void __launch_bounds__(512, 1) k(int iters, long long* out, unsigned* sink) {
  __shared__ uint32_t tmem_ptr;
  if (threadIdx.x < 32) tmem_alloc(&tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  const int warp = threadIdx.x >> 5;
  const uint32_t base = tmem + (((uint32_t)(warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t acc = 0;
  uint32_t r[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) r[j] = threadIdx.x + j;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld_32x32(base + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= r[j];
      }
    } else if (MODE == 1) {
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_st_32x32(base + c * 32, r);
      tmem_st_wait();
    } else {
      uint32_t q[16];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        tmem_ld_32x16(base + c * 16, q);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc ^= q[j];
      }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int MODE>
static void run(const char* name, int warps) {
  long long* d; unsigned* s;
  cudaMalloc(&d, 16); cudaMalloc(&s, 16);
  const int iters = 20000;
  k<MODE><<<148, warps * 32>>>(iters, d, s);
  k<MODE><<<148, warps * 32>>>(iters, d, s);
  cudaError_t e = cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double bytes = (double)iters * 4 * 4096 * warps;       // per SM
  printf("%-16s %2d warps: %7.1f B/clk/SM  (%s)\n", name, warps, bytes / (double)h, cudaGetErrorString(e));
  cudaFree(d); cudaFree(s);
}

int main() {
  for (int w : {4, 8, 16}) run<0>("tcgen05.ld x32", w);
  for (int w : {4, 8, 16}) run<2>("tcgen05.ld x16", w);
  for (int w : {4, 8, 16}) run<1>("tcgen05.st x32", w);
  return 0;
}
