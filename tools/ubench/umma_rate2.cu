// Microbenchmark 2: execution rate of tcgen05.mma kind::f16 (bf16, K=16, both operands in shared memory) as a function of
// the tile shape (cta_group 1 / 2, N) with operands cycling through 6 smem stages like the GEMM ring, and optionally with
// concurrent traffic from other warps of the same CTA: tcgen05.ld of the other accumulator (LOAD & 1), st.shared into an
// unused region (LOAD & 2) — i.e. what the epilogue warps / the TMA unit do while the issuer runs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I panacea_b200/csrc -o tools/ubench/umma_rate2 tools/ubench/umma_rate2.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "ptx.cuh"

using namespace pn;

constexpr int STAGE = 32768, NST = 6;

template <int N, int NCTA, int LOAD>
__global__ void __launch_bounds__(384, 1) rate_kernel(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[4];
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int stop;
  const uint32_t rank = NCTA == 2 ? cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < NST * STAGE / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u; h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
    reinterpret_cast<uint32_t*>(smem)[i] = (h & 0x807f807fu) | 0x3f803f80u;      // two bf16 in [-2, 2)
  }
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    stop = 0;
    mbar_init(&bar, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar2[i], 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) { if (NCTA == 2) tmem_alloc_2sm(&tmem_ptr, 512); else tmem_alloc(&tmem_ptr, 512); }
  tc_fence_before();
  if (NCTA == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = umma_idesc_bf16(128 * NCTA, N, 0, 0);
    const uint64_t dA_ = umma_smem_desc(smem_u32(smem), 16, 1024);
    const uint64_t dB_ = umma_smem_desc(smem_u32(smem) + 16384, 16, 1024);
    const long long t0 = clock64();
    int stage = 0;
    for (int i = 0; i < iters; ++i) {
      tc_fence_after();
      const uint64_t dA = dA_ + (uint64_t)(STAGE >> 4) * stage, dB = dB_ + (uint64_t)(STAGE >> 4) * stage;
      if (++stage == NST) stage = 0;
      const uint32_t d = tmem + ((i / 5) & 1) * (N <= 256 ? N : 0);       // a new accumulator every 5 k-blocks
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (NCTA == 2) umma_f16_ss_2sm(d, dA + 2 * k, dB + 2 * k, idesc, 1u);
        else umma_f16_ss(d, dA + 2 * k, dB + 2 * k, idesc, 1u);
      }
      if (NCTA == 2) umma_commit_2sm(&bar2[i & 3], 3); else umma_commit(&bar2[i & 3]);
    }
    if (NCTA == 2) umma_commit_2sm(&bar, 3); else umma_commit(&bar);
    const long long t1 = clock64();
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    stop = 1;
  } else if (warp >= 4 && rank == 0) {
    // background traffic from 8 warps (TMEM lane quarter = warp % 4)
    uint32_t acc = 0;
    if (LOAD & 1) {
      while (!stop) {
        uint32_t r[32];
        tmem_ld_32x32(tmem + (((uint32_t)(warp & 3) * 32) << 16) + 256 + (warp >= 8 ? 32 : 0), r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= r[j];
      }
    }
    if (LOAD & 2) {
      // the smem ring is 6 x 32 KB; stages' tails [16384 + N/NCTA*128, 32768) are unused when N/NCTA < 128: write elsewhere instead:
      uint4* dst = reinterpret_cast<uint4*>(smem + NST * STAGE) + (threadIdx.x - 128);
      while (!stop) {
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j * 256] = make_uint4(acc, j, 0, 0);       // 8 warps x 512 B x 8 = 32 KB region
        ++acc;
      }
    }
    if (acc == 0x12345678u) out[2] = acc;
  }
  tc_fence_before();
  if (NCTA == 2) {
    if (rank == 1 && threadIdx.x == 0) mbar_wait(&bar, 0);
    cluster_sync_all();
  } else __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); if (NCTA == 2) tmem_dealloc_2sm(tmem, 512); else tmem_dealloc(tmem, 512); }
}

template <int N, int NCTA, int LOAD>
static void run(const char* name) {
  long long* d;
  cudaMalloc(&d, 64);
  const int smem = NST * STAGE + 32768 + 1024;
  auto kern = rate_kernel<N, NCTA, LOAD>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = NCTA; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  const int iters_big = 100000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaLaunchKernelEx(&cfg, kern, 2000, d);
  cudaEventRecord(e0);
  cudaLaunchKernelEx(&cfg, kern, iters_big, d);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const double cyc = (double)h[1] / (iters_big * 4.0);
  const double ns = ms * 1e6 / (iters_big * 4.0);
  const double floor_cyc = (double)(128 * NCTA) * N / (256.0 * NCTA);
  printf("%-34s %7.1f cyc/MMA (floor %5.0f)  %7.2f ns/MMA -> %5.0f MHz  %7.1f TF/s chip (%s)\n", name, cyc, floor_cyc, ns, cyc / ns * 1e3,
         2.0 * 128 * N * 16 * 148 / ns * 1e-3, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<128, 1, 0>("1-CTA N128");
  run<160, 1, 0>("1-CTA N160");
  run<256, 1, 0>("1-CTA N256");
  run<128, 2, 0>("2-CTA N128");
  run<160, 2, 0>("2-CTA N160");
  run<192, 2, 0>("2-CTA N192");
  run<224, 2, 0>("2-CTA N224");
  run<256, 2, 0>("2-CTA N256");
  run<160, 2, 1>("2-CTA N160 + tcgen05.ld traffic");
  run<160, 2, 2>("2-CTA N160 + st.shared traffic");
  run<160, 2, 3>("2-CTA N160 + both");
  run<256, 2, 1>("2-CTA N256 + tcgen05.ld traffic");
  run<256, 2, 2>("2-CTA N256 + st.shared traffic");
  run<128, 1, 1>("1-CTA N128 + tcgen05.ld traffic");
  run<128, 1, 2>("1-CTA N128 + st.shared traffic");
  return 0;
}
