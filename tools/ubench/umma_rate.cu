// Microbenchmark: issue rate of tcgen05.mma (kind::f16, bf16, K=16) from one thread, operands resident in shared
// memory, as a function of N, of the accumulator dependency pattern and of the issue-loop shape.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I panacea_b200/csrc -o tools/ubench/umma_rate tools/ubench/umma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "ptx.cuh"

using namespace pn;

template <int N, int NACC, bool TS, int COMMIT_EVERY = 0, bool FENCE = false, int SPIN = 0>
__global__ void __launch_bounds__(384, 1) rate_kernel(int iters, long long* out, int nstages, int randomize) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[4];
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < 6 * 32768 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (randomize) {
    for (int i = threadIdx.x; i < 6 * 32768 / 4; i += blockDim.x) {
      uint32_t h = (uint32_t)i * 2654435761u; h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
      // two bf16 in [-2, 2): sign | exponent 0x3f/0x40 | random mantissa
      reinterpret_cast<uint32_t*>(smem)[i] = (h & 0x807f807fu) | 0x3f803f80u;
    }
  }
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&bar2[i], 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    const uint64_t dA_ = umma_smem_desc(smem_u32(smem), 16, 1024);
    const uint64_t dB_ = umma_smem_desc(smem_u32(smem) + 16384, 16, 1024);
    const long long t0 = clock64();
    int stage = 0;
    for (int i = 0; i < iters; ++i) {
      if (FENCE) tc_fence_after();
      const uint64_t dA = dA_ + (uint64_t)(nstages * 2048) * stage, dB = dB_ + (uint64_t)(nstages * 2048) * stage;
      if (++stage == 6) stage = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t d = tmem + ((NACC > 1) ? ((k % NACC) * N) : 0);
        if (TS) umma_f16_ts(d, tmem + 384 + 8 * k, dB + 2 * k, idesc, 1u);
        else umma_f16_ss(d, dA + 2 * k, dB + 2 * k, idesc, 1u);
        if (COMMIT_EVERY == 1) umma_commit(&bar2[k]);
      }
      if (COMMIT_EVERY == 4) umma_commit(&bar2[i & 3]);
      if (COMMIT_EVERY == 8 && (i & 1)) umma_commit(&bar2[i & 3]);
    }
    umma_commit(&bar);
    const long long t1 = clock64();
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  if (SPIN && threadIdx.x >= 128) mbar_wait(&bar, 0);      // 8 warps polling an mbarrier like idle epilogue warps
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, int NACC, bool TS, int COMMIT_EVERY = 0, bool FENCE = false, int SPIN = 0>
static void run(const char* name, int grid, int randomize = 0, int nstages = 0) {
  long long* d;
  cudaMalloc(&d, 16);
  const int smem = 6 * 32768 + 1024;
  cudaFuncSetAttribute(rate_kernel<N, NACC, TS, COMMIT_EVERY, FENCE, SPIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  const int iters_big = 200000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  rate_kernel<N, NACC, TS, COMMIT_EVERY, FENCE, SPIN><<<grid, SPIN ? 384 : 128, smem>>>(iters, d, nstages, randomize);
  cudaEventRecord(e0);
  rate_kernel<N, NACC, TS, COMMIT_EVERY, FENCE, SPIN><<<grid, SPIN ? 384 : 128, smem>>>(iters_big, d, nstages, randomize);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const double cyc = (double)h[1] / (iters_big * 4.0);
  const double ns = ms * 1e6 / (iters_big * 4.0);
  printf("%-40s grid=%3d rand=%d  %7.1f cyc/MMA  %7.2f ns/MMA  -> %6.0f MHz  %7.1f TF/s chip (%s)\n", name, grid, randomize, cyc, ns,
         cyc / ns * 1e3, 2.0 * 128 * N * 16 * grid / ns * 1e-3, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  const int grid = 148;
  for (int st = 0; st < 2; ++st) {
    printf("--- %s\n", st ? "6 smem stages cycled (new A and B every 4 MMAs)" : "one smem stage reused");
    run<32, 1, false, 4, true>("SS N32  fence+commit4", grid, 1, st);
    run<64, 1, false, 4, true>("SS N64  fence+commit4", grid, 1, st);
    run<128, 1, false, 4, true>("SS N128 fence+commit4", grid, 1, st);
    run<128, 2, false, 4, true>("SS N128 2 acc fence+commit4", grid, 1, st);
  }
  return 0;
}
