// Microbenchmark: what one thread per SM gets out of the TMA unit, all 148 SMs at once, as a function of the box shape.
//   loads : box = [IB inner bytes] x [ROWS rows] x [ATOMS slabs], slabs IB bytes apart in global memory (the 64-channel
//           swizzle slabs of a GEMM operand), NST boxes in flight;
//   stores: the same boxes written back (bulk-group completion), as the streaming GEMM epilogue does.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I panacea_b200/csrc -o tools/ubench/tma_rate tools/ubench/tma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ptx.cuh"

using namespace pn;

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

template <int NST, bool STORE>
__global__ void __launch_bounds__(128, 1) tma_kernel(const __grid_constant__ CUtensorMap map, int iters, int box_bytes, int rows, int atoms,
                                                     int atoms_total, int row_tiles, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  __shared__ uint64_t bar[NST];
  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) mbar_init(&bar[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    const int kblocks = atoms_total / atoms;
    for (int i = 0; i < iters; ++i) {
      const int slot = i % NST;
      const int tile = (blockIdx.x + gridDim.x * (i / kblocks)) % row_tiles;
      const int kb = i % kblocks;
      if (STORE) {
        tma_store_3d(&map, smem + slot * box_bytes, 0, tile * rows, kb * atoms);
        tma_store_commit();
        tma_store_wait_read_n<NST - 1>();
      } else {
        if (i >= NST) mbar_wait(&bar[slot], ((i / NST) - 1) & 1);
        mbar_arrive_expect_tx(&bar[slot], box_bytes);
        tma_load_3d(smem + slot * box_bytes, &map, &bar[slot], 0, tile * rows, kb * atoms);
      }
    }
    if (STORE) tma_store_wait_all();
    else for (int i = iters; i < iters + NST; ++i) mbar_wait(&bar[i % NST], ((i / NST) - 1) & 1);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
}

// thread 0 streams 2-atom A boxes (32 KB, 3 in flight) while thread 32 stores 64-byte-row chunks (8 KB, 5 in flight)
__global__ void __launch_bounds__(128, 1) both_kernel(const __grid_constant__ CUtensorMap mapL, const __grid_constant__ CUtensorMap mapS,
                                                      int iters, int row_tiles, int do_load, int do_store, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  __shared__ uint64_t bar[3];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) mbar_init(&bar[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0 && do_load) {
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int slot = i % 3;
      if (i >= 3) mbar_wait(&bar[slot], ((i / 3) - 1) & 1);
      mbar_arrive_expect_tx(&bar[slot], 32768);
      tma_load_3d(smem + slot * 32768, &mapL, &bar[slot], 0, ((blockIdx.x + gridDim.x * (i / 2)) % row_tiles) * 128, (i % 2) * 2);
    }
    for (int i = iters; i < iters + 3; ++i) mbar_wait(&bar[i % 3], ((i / 3) - 1) & 1);
    if (blockIdx.x == 0) out[0] = clock64() - t0;
  }
  if (threadIdx.x == 32 && do_store) {
    const long long t0 = clock64();
    const int n = iters * 4;                 // 8 KB chunks: the same bytes as the loads
    for (int i = 0; i < n; ++i) {
      tma_store_3d(&mapS, smem + 98304 + (i % 5) * 8192, 0, ((blockIdx.x + gridDim.x * (i / 5)) % row_tiles) * 128, i % 5);
      tma_store_commit();
      tma_store_wait_read_n<4>();
    }
    tma_store_wait_all();
    if (blockIdx.x == 0) out[1] = clock64() - t0;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn enc;
static void* base;

// matrix [M, C] bf16; a box = IB inner bytes x ROWS rows x ATOMS slabs IB bytes apart
template <int NST, bool STORE>
static void run(int C, int IB, int ROWS, int ATOMS, long long footprint_mb = 100) {
  const int M = (int)(footprint_mb * 1000000 / (C * 2) / 1024 * 1024);
  CUtensorMap map;
  cuuint64_t gdim[3] = {(cuuint64_t)(IB / 2), (cuuint64_t)M, (cuuint64_t)(C * 2 / IB)};
  cuuint64_t gstr[2] = {(cuuint64_t)C * 2, (cuuint64_t)IB};
  cuuint32_t box[3] = {(cuuint32_t)(IB / 2), (cuuint32_t)ROWS, (cuuint32_t)ATOMS}, estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = IB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : IB == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return; }
  long long* d;
  cudaMalloc(&d, 16);
  const int box_bytes = IB * ROWS * ATOMS;
  const int smem = NST * box_bytes + 1024;
  if (smem > 232448) { printf("skip (smem)\n"); return; }
  auto kern = tma_kernel<NST, STORE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int atoms_total = C * 2 / IB;
  const int iters = (int)(4ll * 110000000 / 148 / box_bytes);
  kern<<<148, 128, smem>>>(map, iters, box_bytes, ROWS, ATOMS, atoms_total, M / ROWS, d);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kern<<<148, 128, smem>>>(map, iters, box_bytes, ROWS, ATOMS, atoms_total, M / ROWS, d);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
  long long h;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double cyc = (double)h / iters;
  printf("%s C=%4d (%4lld MB)  box %3d B x %3d rows x %d  = %5.1f KB  x%2d in flight: %7.1f cyc/instr  %5.1f B/clk/SM  chip %6.2f TB/s  (%s)\n",
         STORE ? "store" : "load ", C, footprint_mb, IB, ROWS, ATOMS, box_bytes / 1024.0, NST, cyc, box_bytes / cyc,
         (double)box_bytes * iters * 148 / (ms * 1e-3) * 1e-12, cudaGetErrorString(e));
  cudaFree(d);
}

static void run_both(int do_load, int do_store) {
  const int C = 320, M = 172032;
  CUtensorMap mapL, mapS;
  {
    cuuint64_t gdim[3] = {64, (cuuint64_t)M, (cuuint64_t)(C / 64)};
    cuuint64_t gstr[2] = {(cuuint64_t)C * 2, 128};
    cuuint32_t box[3] = {64, 128, 2}, estr[3] = {1, 1, 1};
    enc(&mapL, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  {
    cuuint64_t gdim[3] = {32, (cuuint64_t)M, (cuuint64_t)(C / 32)};
    cuuint64_t gstr[2] = {(cuuint64_t)C * 2, 64};
    cuuint32_t box[3] = {32, 128, 1}, estr[3] = {1, 1, 1};
    enc(&mapS, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (char*)base + 600ll * 1000000, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  long long* d;
  cudaMalloc(&d, 16);
  cudaMemset(d, 0, 16);
  const int smem = 98304 + 40960 + 1024;
  cudaFuncSetAttribute(both_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 3000;
  both_kernel<<<148, 128, smem>>>(mapL, mapS, iters, M / 128, do_load, do_store, d);
  both_kernel<<<148, 128, smem>>>(mapL, mapS, iters, M / 128, do_load, do_store, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("concurrent load=%d store=%d : load %5.1f B/clk/SM   store %5.1f B/clk/SM  (%s)\n", do_load, do_store,
         do_load ? 32768.0 * iters / h[0] : 0.0, do_store ? 32768.0 * iters / h[1] : 0.0, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no encode fn\n"); return 1; }
  enc = reinterpret_cast<EncodeTiledFn>(p);
  cudaMalloc(&base, 1200ll * 1000000);
  cudaMemset(base, 0, 1200ll * 1000000);
  // loads and stores together: one engine
  run_both(1, 0);
  run_both(0, 1);
  run_both(1, 1);
  // loads, L2-resident source
  run<6, false>(320, 128, 128, 1);
  run<6, false>(320, 128, 64, 1);
  run<6, false>(320, 128, 256, 1);
  run<6, false>(320, 128, 128, 2);
  run<3, false>(320, 128, 128, 2);
  run<4, false>(320, 128, 128, 2);
  run<6, false>(320, 128, 64, 2);
  run<6, false>(1280, 128, 128, 2);
  run<3, false>(1280, 128, 128, 4);
  run<6, false>(320, 64, 128, 1);
  run<6, false>(320, 64, 128, 2);
  run<6, false>(320, 64, 128, 4);
  // loads streamed from DRAM (footprint >> L2)
  run<6, false>(320, 128, 128, 1, 1000);
  run<3, false>(320, 128, 128, 2, 1000);
  run<6, false>(320, 128, 128, 2, 1000);
  // stores
  run<5, true>(320, 64, 128, 1);
  run<5, true>(320, 64, 128, 1, 1000);
  run<2, true>(320, 64, 128, 5);
  run<2, true>(320, 64, 128, 5, 1000);
  run<4, true>(320, 128, 128, 1);
  run<4, true>(320, 128, 128, 1, 1000);
  run<2, true>(320, 128, 128, 2, 1000);
  return 0;
}
