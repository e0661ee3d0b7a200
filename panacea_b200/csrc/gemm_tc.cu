// tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   out[row, n] = epilogue( sum_{tap, c} A[pixel(row) + tap][c] * Wt[n][tap * C + c] )
//
// A is a channels-last bf16 activation tensor [NB, H, W, C] read through ONE rank-4 TMA tensor map; the
// "taps" (1x1 for nn.Linear, 3x3 for the panorama Conv2d, 3x1 over the frame axis for the temporal
// Conv1d) are coordinate offsets of the same map, so im2col never exists in memory and the conv zero
// padding is TMA's out-of-bounds zero fill (outer panorama border only, reference: openaimodel.py:413,
// 455-462 Conv2d(padding=1) on the width-concatenated 6-view image; :418,468-476 Conv1d(k=3,padding=1)).
// B is the packed weight matrix [N, taps*C] (K-major, bf16). Accumulation is fp32 in TMEM.
//
// Kernel structure (persistent, one CTA per SM, 192 threads):
//   warp 0     : TMA producer  (A box [tn,th,tw,64] + B box [BN,64] per k-block, 128B swizzle)
//   warp 1     : TMEM alloc + UMMA issuer (tcgen05.mma cta_group::1 kind::f16, M=128, N=BN, K=16)
//   warps 2..5 : epilogue (tcgen05.ld -> bias/row-vector/GEGLU -> smem transpose -> +residual -> global)
// Pipelines: smem full/empty ring (TMA<->MMA) and a 2-deep TMEM accumulator ring (MMA<->epilogue) so the
// epilogue of tile i overlaps the main loop of tile i+1.
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/panacea_b200.h"


namespace pn {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle atom row
constexpr int GEMM_THREADS = 192;

struct GemmParams {
  CUtensorMap mapA;
  CUtensorMap mapB;
  // geometry of the A tensor / output rows
  int NB, H, W;
  int tw, th, tn;             // tile box extents, tw*th*tn == 128
  int tiles_w, tiles_h, tiles_n, tiles_col;
  int kc_per_tap;             // C / 64
  int taps_h, taps_w, pad_h, pad_w;
  int N;                      // GEMM N (weight rows)
  // epilogue
  void* out;
  const float* bias;
  const float* rowvec;
  const float* residual;
  const float* residual2;
  long long ldo, ldr, ldr2, ldv;
  int rows_per_group, n_groups;
  int out_bf16;
  int geglu;
};

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING_BYTES = 4 * 4096;  // per epilogue warp: 32 rows x 128 B
  static constexpr int ROWMAP_BYTES = 128 * 4;
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + STAGING_BYTES + ROWMAP_BYTES + BAR_BYTES + 1024;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  using S = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint8_t* staging = smem + STAGES * S::STAGE_BYTES;
  int* rowmap = reinterpret_cast<int*>(staging + S::STAGING_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(rowmap) + S::ROWMAP_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_k_blocks = p.taps_h * p.taps_w * p.kc_per_tap;
  const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
  const int num_tiles = tiles_m * p.tiles_col;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static_assert(2 * BN <= 512, "two accumulator stages must fit TMEM");

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.mapA);
    tma_prefetch_desc(&p.mapB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int tcol = tile % p.tiles_col;
        int tm = tile / p.tiles_col;
        const int twi = tm % p.tiles_w; tm /= p.tiles_w;
        const int thi = tm % p.tiles_h; tm /= p.tiles_h;
        const int tni = tm;
        const int x0 = twi * p.tw, y0 = thi * p.th, n0 = tni * p.tn;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          const int tap = kb / p.kc_per_tap;
          const int kc = kb - tap * p.kc_per_tap;
          const int dy = tap / p.taps_w - p.pad_h;
          const int dx = tap % p.taps_w - p.pad_w;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = stage_base + stage * S::STAGE_BYTES;
          uint8_t* sB = sA + S::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          tma_load_4d(sA, &p.mapA, &full_bar[stage], kc * BK, x0 + dx, y0 + dy, n0);
          tma_load_2d(sB, &p.mapB, &full_bar[stage], kb * BK, tcol * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sA = smem_u32(stage_base + stage * S::STAGE_BYTES);
          const uint32_t sB = sA + S::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = umma_smem_desc(sA + k * 32, 16, 1024);
            const uint64_t db = umma_smem_desc(sB + k * 32, 16, 1024);
            umma_f16_ss(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                       // frees the smem slot when the MMAs retire
          if (kb == num_k_blocks - 1) umma_commit(&tmem_full[acc]);  // accumulator ready for the epilogue
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;               // staging buffer index
    const int lane_grp = warp & 3;         // TMEM lane quarter this warp may access
    uint8_t* my_stage = staging + ew * 4096;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int tcol = tile % p.tiles_col;
      int tm = tile / p.tiles_col;
      const int twi = tm % p.tiles_w; tm /= p.tiles_w;
      const int thi = tm % p.tiles_h; tm /= p.tiles_h;
      const int tni = tm;
      // output row of the tile row this thread owns (tile rows are ordered [tn][th][tw] = TMA box order)
      {
        const int r = lane_grp * 32 + lane;
        const int dx = r % p.tw;
        const int dy = (r / p.tw) % p.th;
        const int dn = r / (p.tw * p.th);
        const int x = twi * p.tw + dx, y = thi * p.th + dy, n = tni * p.tn + dn;
        const bool ok = (x < p.W) && (y < p.H) && (n < p.NB);
        rowmap[r] = ok ? ((n * p.H + y) * p.W + x) : -1;
      }
      __syncwarp();
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (uint32_t(lane_grp * 32) << 16) + acc * BN;
      const int my_row = rowmap[lane_grp * 32 + lane];
      const int n_base = tcol * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(t_row + c * 32, v);
        tmem_ld_wait();
        if (c == BN / 32 - 1) {
          // all TMEM reads of this accumulator stage are done -> hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        const int n0 = n_base + c * 32;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + j < p.N) f[j] += __ldg(p.bias + n0 + j);
        }
        if (p.rowvec != nullptr && my_row >= 0) {
          const float* rv = p.rowvec + (long long)((my_row / p.rows_per_group) % p.n_groups) * p.ldv;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + j < p.N) f[j] += __ldg(rv + n0 + j);
        }
        if (p.geglu) {
          // packed weight rows are interleaved (value, gate) pairs: out[n/2] = value * gelu(gate)
          // (reference GEGLU: attention.py:97-99, chunk order value-first, exact erf GELU)
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = f[2 * j] * gelu_erf(f[2 * j + 1]);
          // 16 bf16 = 32 B per row
          uint4* dst = reinterpret_cast<uint4*>(my_stage + lane * 32);
          dst[0] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
          dst[1] = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]), pack_bf16x2(f[14], f[15]));
          __syncwarp();
          __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
          const int no0 = n0 / 2;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int rr = i * 16 + (lane >> 1);
            const int ch = lane & 1;
            const int grow = rowmap[lane_grp * 32 + rr];
            if (grow >= 0 && no0 + ch * 8 < p.N / 2) {
              const uint4 val = *reinterpret_cast<const uint4*>(my_stage + rr * 32 + ch * 16);
              *reinterpret_cast<uint4*>(out + (long long)grow * p.ldo + no0 + ch * 8) = val;
            }
          }
          __syncwarp();
        } else if (p.out_bf16) {
          // 32 bf16 = 64 B per row, 16 B chunks XOR-swizzled with (row>>1)&3
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const int sw = ch ^ ((lane >> 1) & 3);
            *reinterpret_cast<uint4*>(my_stage + lane * 64 + sw * 16) =
                make_uint4(pack_bf16x2(f[ch * 8 + 0], f[ch * 8 + 1]), pack_bf16x2(f[ch * 8 + 2], f[ch * 8 + 3]),
                           pack_bf16x2(f[ch * 8 + 4], f[ch * 8 + 5]), pack_bf16x2(f[ch * 8 + 6], f[ch * 8 + 7]));
          }
          __syncwarp();
          __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + (lane >> 2);
            const int ch = lane & 3;
            const int grow = rowmap[lane_grp * 32 + rr];
            if (grow >= 0 && n0 + ch * 8 < p.N) {
              const int sw = ch ^ ((rr >> 1) & 3);
              uint4 val = *reinterpret_cast<const uint4*>(my_stage + rr * 64 + sw * 16);
              if (p.residual != nullptr) {
                const float* rp = p.residual + (long long)grow * p.ldr + n0 + ch * 8;
                const float4 r0 = *reinterpret_cast<const float4*>(rp);
                const float4 r1 = *reinterpret_cast<const float4*>(rp + 4);
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&val);
                float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
                float2 c2 = __bfloat1622float2(h[2]), d = __bfloat1622float2(h[3]);
                val = make_uint4(pack_bf16x2(a.x + r0.x, a.y + r0.y), pack_bf16x2(b.x + r0.z, b.y + r0.w),
                                 pack_bf16x2(c2.x + r1.x, c2.y + r1.y), pack_bf16x2(d.x + r1.z, d.y + r1.w));
              }
              *reinterpret_cast<uint4*>(out + (long long)grow * p.ldo + n0 + ch * 8) = val;
            }
          }
          __syncwarp();
        } else {
          // 32 fp32 = 128 B per row, 16 B chunks XOR-swizzled with row&7 (conflict-free both ways)
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            const int sw = ch ^ (lane & 7);
            *reinterpret_cast<float4*>(my_stage + lane * 128 + sw * 16) =
                make_float4(f[ch * 4 + 0], f[ch * 4 + 1], f[ch * 4 + 2], f[ch * 4 + 3]);
          }
          __syncwarp();
          float* out = reinterpret_cast<float*>(p.out);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + (lane >> 3);
            const int ch = lane & 7;
            const int grow = rowmap[lane_grp * 32 + rr];
            if (grow >= 0 && n0 + ch * 4 < p.N) {
              const int sw = ch ^ (rr & 7);
              float4 val = *reinterpret_cast<const float4*>(my_stage + rr * 128 + sw * 16);
              if (p.residual != nullptr) {
                const float4 r = *reinterpret_cast<const float4*>(p.residual + (long long)grow * p.ldr + n0 + ch * 4);
                val.x += r.x; val.y += r.y; val.z += r.z; val.w += r.w;
              }
              if (p.residual2 != nullptr) {
                const float4 r = *reinterpret_cast<const float4*>(p.residual2 + (long long)grow * p.ldr2 + n0 + ch * 4);
                val.x += r.x; val.y += r.y; val.z += r.z; val.w += r.w;
              }
              *reinterpret_cast<float4*>(out + (long long)grow * p.ldo + n0 + ch * 4) = val;
            }
          }
          __syncwarp();
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// Pick the [tn, th, tw] box (product 128, powers of two) that wastes the fewest MMA rows.
static void pick_tile(long long NB, long long H, long long W, int* tw, int* th, int* tn) {
  double best = 1e30;
  for (int a = 128; a >= 1; a >>= 1) {         // tw
    for (int b = 128 / a; b >= 1; b >>= 1) {   // th
      const int c = 128 / (a * b);             // tn
      const long long tiles = ((W + a - 1) / a) * ((H + b - 1) / b) * ((NB + c - 1) / c);
      const double waste = double(tiles) * 128.0 / double(NB * H * W);
      // prefer less waste; among equals prefer the widest inner run
      const double score = waste - 1e-6 * a - 1e-9 * b;
      if (score < best) { best = score; *tw = a; *th = b; *tn = c; }
    }
  }
}

template <int BN, int STAGES>
static int launch_gemm(const GemmParams& p, int num_tiles, cudaStream_t stream) {
  using S = GemmSmem<BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    PN_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    attr_set = true;
  }
  int grid = sm_count();
  if (grid > num_tiles) grid = num_tiles;
  gemm_tc_kernel<BN, STAGES><<<grid, GEMM_THREADS, S::TOTAL, stream>>>(p);
  PN_CHECK_CUDA(cudaGetLastError());
  return PN_OK;
}

}  // namespace pn

using namespace pn;

extern "C" int pn_gemm(const pn_gemm_args* a, void* stream_v) {
  if (a == nullptr) return fail(PN_ERR_INVALID, "pn_gemm: null args");
  PN_REQUIRE(a->A && a->B && a->out, "pn_gemm: null tensor pointer");
  PN_REQUIRE(a->C > 0 && a->C % 64 == 0, "pn_gemm: C=%lld must be a positive multiple of 64", (long long)a->C);
  PN_REQUIRE(a->N > 0 && a->N % 8 == 0, "pn_gemm: N=%d must be a positive multiple of 8", a->N);
  PN_REQUIRE(a->taps_h >= 1 && a->taps_h <= 3 && a->taps_w >= 1 && a->taps_w <= 3, "pn_gemm: taps must be 1..3");
  PN_REQUIRE(a->NB > 0 && a->H > 0 && a->W > 0, "pn_gemm: empty A geometry");
  PN_REQUIRE((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->B) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(a->out) & 15) == 0,
             "pn_gemm: pointers must be 16-byte aligned");
  PN_REQUIRE(a->a_stride_w % 8 == 0 && a->a_stride_h % 8 == 0 && a->a_stride_n % 8 == 0,
             "pn_gemm: A strides must be multiples of 8 elements (16 bytes)");
  const int n_out = a->geglu ? a->N / 2 : a->N;
  PN_REQUIRE(a->ldo >= n_out && a->ldo % 8 == 0, "pn_gemm: ldo=%lld too small or misaligned", (long long)a->ldo);
  if (a->geglu) PN_REQUIRE(a->out_bf16 && a->residual == nullptr && a->N % 32 == 0, "pn_gemm: GEGLU needs bf16 out, no residual, N%%32==0");
  if (a->residual) PN_REQUIRE(a->ldr >= a->N && a->ldr % 4 == 0, "pn_gemm: bad ldr");
  if (a->residual2) PN_REQUIRE(!a->out_bf16 && !a->geglu && a->ldr2 >= a->N && a->ldr2 % 4 == 0, "pn_gemm: residual2 needs fp32 out and a valid ldr2");
  if (a->rowvec) PN_REQUIRE(a->rows_per_group > 0 && a->n_groups > 0, "pn_gemm: rowvec needs rows_per_group/n_groups");

  GemmParams p;
  std::memset(&p, 0, sizeof(p));
  long long NB = a->NB, H = a->H, W = a->W;
  long long sw = a->a_stride_w, sh = a->a_stride_h, sn = a->a_stride_n;
  const bool pointwise = (a->taps_h == 1 && a->taps_w == 1);
  if (pointwise && sh == sw * W && sn == sh * H) {  // plain GEMM over a dense row set: flatten to [1,1,M]
    W = NB * H * W; H = 1; NB = 1; sh = sw * W; sn = sh;
  }
  int tw, th, tn;
  pick_tile(NB, H, W, &tw, &th, &tn);
  p.NB = (int)NB; p.H = (int)H; p.W = (int)W;
  p.tw = tw; p.th = th; p.tn = tn;
  p.tiles_w = (int)((W + tw - 1) / tw);
  p.tiles_h = (int)((H + th - 1) / th);
  p.tiles_n = (int)((NB + tn - 1) / tn);
  p.kc_per_tap = (int)(a->C / 64);
  p.taps_h = a->taps_h; p.taps_w = a->taps_w;
  p.pad_h = a->taps_h / 2; p.pad_w = a->taps_w / 2;
  p.N = a->N;
  p.out = a->out; p.bias = a->bias; p.rowvec = a->rowvec; p.residual = a->residual;
  p.residual2 = a->residual2;
  p.ldo = a->ldo; p.ldr = a->ldr; p.ldr2 = a->ldr2;
  p.ldv = a->rowvec_ld > 0 ? a->rowvec_ld : a->N;
  p.rows_per_group = a->rows_per_group > 0 ? a->rows_per_group : 1;
  p.n_groups = a->n_groups > 0 ? a->n_groups : 1;
  p.out_bf16 = a->out_bf16; p.geglu = a->geglu;

  // N tile: every channel count of the network is a multiple of 160 (320/640/960/1280/1920/2560/5120);
  // fall back to 128 / 64 / 32 otherwise.
  int BN;
  if (a->N % 160 == 0) BN = 160;
  else if (a->N >= 128) BN = 128;
  else if (a->N > 32) BN = 64;
  else BN = 32;
  p.tiles_col = (a->N + BN - 1) / BN;

  const uint64_t dimsA[4] = {(uint64_t)a->C, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
  const uint64_t strA[3] = {(uint64_t)sw, (uint64_t)sh, (uint64_t)sn};
  const uint32_t boxA[4] = {64u, (uint32_t)tw, (uint32_t)th, (uint32_t)tn};
  int rc = cached_tmap_bf16(&p.mapA, a->A, 4, dimsA, strA, boxA, 128);
  if (rc != PN_OK) return rc;
  const uint64_t K = (uint64_t)a->taps_h * a->taps_w * a->C;
  const uint64_t dimsB[2] = {K, (uint64_t)a->N};
  const uint64_t strB[1] = {K};
  const uint32_t boxB[2] = {64u, (uint32_t)BN};
  rc = cached_tmap_bf16(&p.mapB, a->B, 2, dimsB, strB, boxB, 128);
  if (rc != PN_OK) return rc;

  const int num_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.tiles_col;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  switch (BN) {
    case 160: return launch_gemm<160, 5>(p, num_tiles, stream);
    case 128: return launch_gemm<128, 6>(p, num_tiles, stream);
    case 64: return launch_gemm<64, 8>(p, num_tiles, stream);
    default: return launch_gemm<32, 8>(p, num_tiles, stream);
  }
}
