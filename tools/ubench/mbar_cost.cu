// Microbenchmark: cost of one mbarrier try_wait on an already-completed phase, of arrive.expect_tx, and of a
// wait -> expect_tx -> TMA-free loop, from a single thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I panacea_b200/csrc -o tools/ubench/mbar_cost tools/ubench/mbar_cost.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace pn;

__global__ void k(long long* out, int iters) {
  __shared__ uint64_t bar[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1);
    fence_barrier_init();
    for (int i = 0; i < 8; ++i) mbar_arrive(&bar[i]);      // phase 0 complete
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) mbar_wait(&bar[i & 7], 0);
    long long t1 = clock64();
    out[0] = t1 - t0;
    // arrive + wait ping-pong on one barrier (phase flips every iteration)
    uint32_t ph = 1;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { mbar_arrive(&bar[0]); mbar_wait(&bar[0], ph); ph ^= 1; }
    t1 = clock64();
    out[1] = t1 - t0;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { mbar_arrive(&bar[1]); }
    t1 = clock64();
    out[2] = t1 - t0;
  }
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  const int iters = 10000;
  k<<<1, 32>>>(d, iters); k<<<1, 32>>>(d, iters);
  cudaDeviceSynchronize();
  long long h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
  printf("try_wait (complete phase): %.1f cyc   arrive+wait: %.1f cyc   arrive: %.1f cyc  (%s)\n", (double)h[0] / iters, (double)h[1] / iters,
         (double)h[2] / iters, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
