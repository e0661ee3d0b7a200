// tcgen05 flash attention for the decomposed-4D attention of Panacea (head_dim 64, bf16 operands, fp32 softmax).
//
// One kernel serves the three tensor-core attention variants; they differ only in which K/V tiles a query tile
// visits, and every tile is a TMA box of ONE rank-5 tensor map over the token buffer [F, H, V, w, C]
// (frames, latent rows, views, columns per view, channels) — the views are never sliced or copied:
//   * intra-view  (attention.py:407-489)  : K/V = the query's own view;
//   * cross-view  (attention.py:518-610)  : K/V = the neighbour views from a table
//                                            {5,1},{0,2},{1,3},{2,4},{3,5},{4} (the reference's asymmetric ring);
//   * text cross-attention (attention.py:229-291, 77 keys): V=1, one K/V block per batch element, tail masked.
// (Temporal self-attention over T<=16 frames is a CUDA-core kernel, attn_small.cu.)
//
// CTA = one query tile (<=128 queries of one frame/view/head). Warp roles:
//   warp 0: TMA producer (Q once; K and V boxes through a 3-stage ring)
//   warp 1: UMMA issuer   S = Q K^T  (M=128, N=kv_n, K=64)  -> TMEM, double buffered
//                         PV = P V   (M=128, N=64,  K=kv_n) -> TMEM, double buffered (V is the MN-major B operand)
//   warps 2-9: softmax, two threads per query row (TMEM lane == row; the pair splits the S columns and the
//              output channels): tcgen05.ld S once -> running max (pair exchange through smem) -> exp2 -> bf16 P
//              into 128B-swizzled smem (A operand of PV); O accumulates in registers with the usual rescale; the
//              PV of block j-1 is folded in while the tensor core works on block j.
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/panacea_b200.h"

namespace pn {

constexpr int FA_D = 64;
constexpr int FA_STAGES = 3;
constexpr int FA_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 softmax
constexpr int FA_TILE_BYTES = 128 * 128;          // 128 rows x 64 bf16
constexpr int FA_SMEM_Q = 0;                                        // 2 query tiles (double buffered across tiles)
constexpr int FA_SMEM_K = 2 * FA_TILE_BYTES;
constexpr int FA_SMEM_V = FA_SMEM_K + FA_STAGES * FA_TILE_BYTES;
constexpr int FA_SMEM_P = FA_SMEM_V + FA_STAGES * FA_TILE_BYTES;   // 2 buffers x 2 atoms x 16 KB
constexpr int FA_SMEM_BAR = FA_SMEM_P + 4 * FA_TILE_BYTES;
constexpr int FA_SMEM_XCH = FA_SMEM_BAR + 256;                      // row-max exchange [2][2][128] + row-sum exchange [2][128]
constexpr int FA_SMEM_TOTAL = FA_SMEM_XCH + (2 * 2 * 128 + 2 * 128) * 4 + 1024;

struct FaParams {
  CUtensorMap mapQ;
  CUtensorMap mapK;
  CUtensorMap mapV;
  int heads;
  int F, H, V, W;              // query token grid
  int qw, qh, tiles_x, tiles_y;
  int kw, kh, kv_rows, kv_n, kv_yblocks;
  int kv_views[8][2];
  int kv_view_count[8];
  int kv_frame_div;            // kv frame = q frame / kv_frame_div
  int total_tiles;
  float scale_log2;            // softmax scale * log2(e)
  __nv_bfloat16* out;
  long long out_ld;            // token stride of out (elements)
};

struct FaTile {
  int x0, y0, head, view, frame, nblk;
};

// q tile fastest, then head, view, frame: CTAs that run concurrently share K/V in L2
__device__ __forceinline__ FaTile fa_decode(const FaParams& p, int tile) {
  FaTile t;
  const int tx = tile % p.tiles_x; tile /= p.tiles_x;
  const int ty = tile % p.tiles_y; tile /= p.tiles_y;
  t.head = tile % p.heads; tile /= p.heads;
  t.view = tile % p.V; tile /= p.V;
  t.frame = tile;
  t.x0 = tx * p.qw;
  t.y0 = ty * p.qh;
  t.nblk = p.kv_view_count[t.view] * p.kv_yblocks;
  return t;
}

// Persistent kernel: each CTA walks a strided list of query tiles; the K/V-block pipeline (S and PV double buffers,
// 3-stage K/V ring) runs ACROSS tile boundaries, so the prologue, the Q load and the last PV of a tile hide behind the
// next tile's work (this is what makes the 1-block text attention and the 16-block view attention share one kernel).
// MASK: the key block has padding columns (kv_rows < kv_n, e.g. 77 text keys in an 80-wide block) that must get p = 0.
// NCH: number of 16-column chunks of a key block (kv_n / 16) fixed at compile time for the shapes of the network
// (7 = 112 keys per block at 32x56 views, 8 = 128 keys at 32x64 views, 5 = the 77 text keys, 2 = the 4x7 middle block);
// 0 = read it from the parameters (any other shape).
template <bool MASK, int NCH>
__global__ void __launch_bounds__(FA_THREADS, 1) attn_fa_kernel(const __grid_constant__ FaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FA_SMEM_BAR);
  uint64_t* q_full = bars;                     // [2]
  uint64_t* q_empty = bars + 2;                // [2]
  uint64_t* k_full = bars + 4;                 // [3]
  uint64_t* v_full = bars + 7;                 // [3]
  uint64_t* kv_empty = bars + 10;              // [3]
  uint64_t* s_full = bars + 13;                // [2]
  uint64_t* s_empty = bars + 15;               // [2]
  uint64_t* p_full = bars + 17;                // [2]
  uint64_t* pv_full = bars + 19;               // [2]
  uint64_t* pv_empty = bars + 21;              // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 23);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // zero Q/K/V staging once: rows a TMA box does not cover (kv_rows..kv_n) must read as 0, never as stale NaNs
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < FA_SMEM_P / 16; i += FA_THREADS) z[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.mapQ);
    tma_prefetch_desc(&p.mapK);
    tma_prefetch_desc(&p.mapV);
    for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    for (int i = 0; i < FA_STAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 8);      // one elected arrive per softmax warp
      mbar_init(&p_full[i], 8);
      mbar_init(&pv_full[i], 1);
      mbar_init(&pv_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tm_S = tmem_base;          // 2 x 128 columns
  const uint32_t tm_PV = tmem_base + 256;   // 2 x 64 columns

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t q_bytes = (uint32_t)(p.qw * p.qh) * 128u;
      const uint32_t kv_bytes = (uint32_t)p.kv_rows * 128u;
      int g = 0, it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const FaTile t = fa_decode(p, tile);
        const int qb = it & 1;
        mbar_wait(&q_empty[qb], (uint32_t)(((it >> 1) & 1) ^ 1));
        mbar_arrive_expect_tx(&q_full[qb], q_bytes);
        tma_load_5d(smem + FA_SMEM_Q + qb * FA_TILE_BYTES, &p.mapQ, &q_full[qb], t.head * FA_D, t.x0, t.view, t.y0, t.frame);
        const int kv_frame = t.frame / p.kv_frame_div;
        for (int j = 0; j < t.nblk; ++j, ++g) {
          const int st = g % FA_STAGES;
          const int vi = j / p.kv_yblocks, yb = j - vi * p.kv_yblocks;
          const int kvv = p.kv_views[t.view][vi];
          mbar_wait(&kv_empty[st], (uint32_t)(((g / FA_STAGES) & 1) ^ 1));
          mbar_arrive_expect_tx(&k_full[st], kv_bytes);
          tma_load_5d(smem + FA_SMEM_K + st * FA_TILE_BYTES, &p.mapK, &k_full[st], t.head * FA_D, 0, kvv, yb * p.kh, kv_frame);
          mbar_arrive_expect_tx(&v_full[st], kv_bytes);
          tma_load_5d(smem + FA_SMEM_V + st * FA_TILE_BYTES, &p.mapV, &v_full[st], t.head * FA_D, 0, kvv, yb * p.kh, kv_frame);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer (one thread runs the whole loop; descriptors advance by integer adds) ====
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(128, p.kv_n, 0, 0);      // S = Q K^T : both K-major
      const uint32_t idesc_pv = umma_idesc_bf16(128, FA_D, 0, 1);       // PV: A = P K-major, B = V MN-major
      const int ksteps_pv = p.kv_n / 16;
      const uint64_t dQ0 = umma_smem_desc(smem_u32(smem + FA_SMEM_Q), 16, 1024);
      const uint64_t dK0 = umma_smem_desc(smem_u32(smem + FA_SMEM_K), 16, 1024);
      const uint64_t dP0 = umma_smem_desc(smem_u32(smem + FA_SMEM_P), 16, 1024);
      const uint64_t dV0 = umma_smem_desc(smem_u32(smem + FA_SMEM_V), 1024, 1024);
      constexpr uint64_t TILE_STEP = FA_TILE_BYTES >> 4;                // start-address field is in 16-byte units
      auto issue_pv = [&](int i) {
        const int st = i % FA_STAGES, buf = i & 1;
        mbar_wait(&v_full[st], (uint32_t)((i / FA_STAGES) & 1));
        mbar_wait(&p_full[buf], (uint32_t)((i >> 1) & 1));
        mbar_wait(&pv_empty[buf], (uint32_t)(((i >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint64_t dP = dP0 + TILE_STEP * (2 * buf);
        const uint64_t dV = dV0 + TILE_STEP * st;
        for (int k = 0; k < ksteps_pv; ++k) {
          // P: 64-key tiles of 16 KB, 32 B per K step inside one; V: 16 keys (rows of 128 B) per K step
          umma_f16_ss(tm_PV + buf * FA_D, dP + TILE_STEP * (k >> 2) + 2 * (k & 3), dV + 128 * k, idesc_pv,
                      k > 0 ? 1u : 0u);
        }
        umma_commit(&pv_full[buf]);
        umma_commit(&kv_empty[st]);
      };
      int g = 0, it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const FaTile t = fa_decode(p, tile);
        const int qb = it & 1;
        const uint64_t dQ = dQ0 + TILE_STEP * qb;
        mbar_wait(&q_full[qb], (uint32_t)((it >> 1) & 1));
        for (int j = 0; j < t.nblk; ++j, ++g) {
          const int st = g % FA_STAGES, buf = g & 1;
          mbar_wait(&k_full[st], (uint32_t)((g / FA_STAGES) & 1));
          mbar_wait(&s_empty[buf], (uint32_t)(((g >> 1) & 1) ^ 1));
          tc_fence_after();
          const uint64_t dK = dK0 + TILE_STEP * st;
#pragma unroll
          for (int k = 0; k < FA_D / 16; ++k)
            umma_f16_ss(tm_S + buf * 128, dQ + 2 * k, dK + 2 * k, idesc_s, k > 0 ? 1u : 0u);
          umma_commit(&s_full[buf]);
          if (j == t.nblk - 1) umma_commit(&q_empty[qb]);   // every S MMA reading this Q tile has retired
          if (g > 0) issue_pv(g - 1);
        }
      }
      if (g > 0) issue_pv(g - 1);
    }
    __syncwarp();
  } else {
    // ===================== softmax / output warps =====================
    // Two threads per query row: warps w and w+4 share a TMEM lane quarter; the first takes the even 16-column
    // chunks of S and output channels [0,32), the second the odd chunks and channels [32,64). Each reads its S
    // values from TMEM once (registers), the pair agrees on the running row maximum through shared memory (one
    // 64-thread named barrier per block), row sums are combined once per tile.
    const int sw_id = warp - 2;
    const int lane_grp = warp & 3;
    const int half = sw_id >> 2;
    const int row = lane_grp * 32 + lane;
    const uint32_t lane_addr = uint32_t(lane_grp * 32) << 16;
    float* xch = reinterpret_cast<float*>(smem + FA_SMEM_XCH);      // [buf][half][row]
    float* xch_l = xch + 2 * 2 * 128;                               // [half][row]
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;
    float O[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) O[i] = 0.f;
    const int nchunk = NCH > 0 ? NCH : p.kv_n / 16;
    const float c = p.scale_log2;
    const uint32_t bar_id = 1 + lane_grp;

    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory"); };

    auto consume_pv = [&](int i, float alpha) {
      const int buf = i & 1;
      mbar_wait(&pv_full[buf], (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32(tm_PV + lane_addr + buf * FA_D + half * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int t = 0; t < 32; ++t) O[t] = O[t] * alpha + __uint_as_float(v[t]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pv_empty[buf]);
    };

    // combine the pair's partial row sums, normalise and store this thread's 32 channels; then clear O
    auto finish_tile = [&](const FaTile& t, float l_part) {
      pair_sync();
      xch_l[half * 128 + row] = l_part;
      pair_sync();
      const float l_tot = l_part + xch_l[(half ^ 1) * 128 + row];
      const int yy = row / p.qw, xx = row - yy * p.qw;
      const int x = t.x0 + xx, y = t.y0 + yy;
      if (row < p.qw * p.qh && x < p.W && y < p.H) {
        const float inv = 1.f / l_tot;
        const long long token = (((long long)t.frame * p.H + y) * p.V + t.view) * p.W + x;
        __nv_bfloat16* dst = p.out + token * p.out_ld + t.head * FA_D + half * 32;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          *reinterpret_cast<uint4*>(dst + i * 8) =
              make_uint4(pack_bf16x2(O[i * 8 + 0] * inv, O[i * 8 + 1] * inv), pack_bf16x2(O[i * 8 + 2] * inv, O[i * 8 + 3] * inv),
                         pack_bf16x2(O[i * 8 + 4] * inv, O[i * 8 + 5] * inv), pack_bf16x2(O[i * 8 + 6] * inv, O[i * 8 + 7] * inv));
        }
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) O[i] = 0.f;
    };

    int g = 0;
    FaTile prev;
    prev.x0 = prev.y0 = prev.head = prev.view = prev.frame = prev.nblk = 0;
    float l_prev = 0.f;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const FaTile t = fa_decode(p, tile);
      l_prev = l_run;
      m_run = -INFINITY;
      l_run = 0.f;
      for (int j = 0; j < t.nblk; ++j, ++g) {
        const int buf = g & 1;
        mbar_wait(&s_full[buf], (uint32_t)((g >> 1) & 1));
        tc_fence_after();
        const uint32_t tS = tm_S + lane_addr + buf * 128;
        // my chunks of S -> registers (single TMEM pass)
        uint32_t sv[4][16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int ch = half + 2 * q;
          if (ch < nchunk) tmem_ld_32x16(tS + ch * 16, sv[q]);
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[buf]);  // S buffer may be overwritten by the MMA of block g+2
        // four independent max chains per chunk (a single 56-long dependent chain would cost ~4 cycles per element)
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int ch = half + 2 * q;
          if (ch < nchunk) {
#pragma unroll
            for (int tt = 0; tt < 16; ++tt)
              if (!MASK || ch * 16 + tt < p.kv_rows) mx4[tt & 3] = fmaxf(mx4[tt & 3], __uint_as_float(sv[q][tt]));
          }
        }
        float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        float* xb = xch + buf * 256;
        xb[half * 128 + row] = mx;
        pair_sync();
        mx = fmaxf(mx, xb[(half ^ 1) * 128 + row]);
        const float m_new = fmaxf(m_run, mx * c);
        const float alpha = (m_run == -INFINITY) ? 0.f : ex2_approx(m_run - m_new);
        float rs4[4] = {0.f, 0.f, 0.f, 0.f};
        uint8_t* sP = smem + FA_SMEM_P + buf * 2 * FA_TILE_BYTES;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int ch = half + 2 * q;
          if (ch < nchunk) {
            float e[16];
#pragma unroll
            for (int tt = 0; tt < 16; ++tt) {
              const float pv = ex2_approx(__uint_as_float(sv[q][tt]) * c - m_new);
              e[tt] = (!MASK || ch * 16 + tt < p.kv_rows) ? pv : 0.f;
              rs4[tt & 3] += e[tt];
            }
            // 16 keys = 2 chunks of 16 B inside atom (ch/4); chunk index within the 128 B row = (ch%4)*2 + {0,1}
            uint8_t* atom = sP + (ch >> 2) * FA_TILE_BYTES + row * 128;
            const int c0 = (ch & 3) * 2;
            *reinterpret_cast<uint4*>(atom + ((c0 ^ (row & 7)) << 4)) =
                make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
            *reinterpret_cast<uint4*>(atom + (((c0 + 1) ^ (row & 7)) << 4)) =
                make_uint4(pack_bf16x2(e[8], e[9]), pack_bf16x2(e[10], e[11]), pack_bf16x2(e[12], e[13]), pack_bf16x2(e[14], e[15]));
          }
        }
        fence_proxy_async_smem();                   // every writer publishes its P stores to the async proxy ...
        __syncwarp();                               // ... before the warp's single arrive
        if (lane == 0) mbar_arrive(&p_full[buf]);
        l_run = l_run * alpha + ((rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
        m_run = m_new;
        if (g > 0) {
          consume_pv(g - 1, alpha_prev);            // block g-1 may be the last block of the previous tile
          if (j == 0) finish_tile(prev, l_prev);
        }
        alpha_prev = alpha;
      }
      prev = t;
    }
    if (g > 0) {
      consume_pv(g - 1, alpha_prev);
      finish_tile(prev, l_run);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

typedef void (*FaKernel)(const FaParams);
static FaKernel fa_kernels[10] = {
    attn_fa_kernel<false, 0>, attn_fa_kernel<false, 2>, attn_fa_kernel<false, 5>, attn_fa_kernel<false, 7>, attn_fa_kernel<false, 8>,
    attn_fa_kernel<true, 0>,  attn_fa_kernel<true, 2>,  attn_fa_kernel<true, 5>,  attn_fa_kernel<true, 7>,  attn_fa_kernel<true, 8>};

}  // namespace pn

using namespace pn;

extern "C" int pn_attention(const pn_attn_args* a, void* stream_v) {
  if (a == nullptr) return fail(PN_ERR_INVALID, "pn_attention: null args");
  PN_REQUIRE(a->q && a->k && a->v && a->out, "pn_attention: null tensor pointer");
  PN_REQUIRE(a->head_dim == FA_D, "pn_attention: head_dim %d unsupported (64 only)", a->head_dim);
  PN_REQUIRE(a->heads > 0 && a->F > 0 && a->H > 0 && a->V > 0 && a->V <= 8 && a->W > 0, "pn_attention: bad query geometry");
  PN_REQUIRE(a->Hk > 0 && a->Vk > 0 && a->Vk <= 8 && a->Wk > 0 && a->kv_frame_div > 0, "pn_attention: bad key geometry");
  PN_REQUIRE(a->q_ld % 8 == 0 && a->kv_ld % 8 == 0 && a->out_ld % 8 == 0, "pn_attention: token strides must be multiples of 8");
  PN_REQUIRE(a->q_ld >= a->heads * FA_D && a->kv_ld >= a->heads * FA_D && a->out_ld >= a->heads * FA_D,
             "pn_attention: token stride smaller than heads*64");

  FaParams p;
  std::memset(&p, 0, sizeof(p));
  p.heads = a->heads;
  p.F = (int)a->F; p.H = (int)a->H; p.V = (int)a->V; p.W = (int)a->W;
  // query tile: full view width when it fits, as many rows as keep <= 128 queries
  p.qw = (int)(a->W <= 128 ? a->W : 128);
  p.qh = 128 / p.qw;
  if (p.qh > a->H) p.qh = (int)a->H;
  if (p.qh < 1) p.qh = 1;
  p.tiles_x = (int)((a->W + p.qw - 1) / p.qw);
  p.tiles_y = (int)((a->H + p.qh - 1) / p.qh);
  // key block: full key-view width (must fit one block row-wise), rows = largest divisor of Hk with <= 128 keys
  PN_REQUIRE(a->Wk <= 128, "pn_attention: key view width %lld > 128 unsupported", (long long)a->Wk);
  p.kw = (int)a->Wk;
  int kh = 128 / p.kw;
  if (kh > a->Hk) kh = (int)a->Hk;
  while (kh > 1 && (a->Hk % kh) != 0) --kh;
  p.kh = kh;
  p.kv_rows = p.kw * p.kh;
  p.kv_n = (p.kv_rows + 15) / 16 * 16;
  p.kv_yblocks = (int)(a->Hk / p.kh);
  for (int v = 0; v < a->V; ++v) {
    const int cnt = a->kv_view_count[v];
    PN_REQUIRE(cnt >= 1 && cnt <= 2, "pn_attention: kv_view_count[%d]=%d must be 1 or 2", v, cnt);
    p.kv_view_count[v] = cnt;
    for (int i = 0; i < cnt; ++i) {
      PN_REQUIRE(a->kv_views[v][i] >= 0 && a->kv_views[v][i] < a->Vk, "pn_attention: kv view out of range");
      p.kv_views[v][i] = a->kv_views[v][i];
    }
  }
  p.kv_frame_div = a->kv_frame_div;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  p.out_ld = a->out_ld;

  const uint64_t chq = (uint64_t)a->heads * FA_D;
  {
    const uint64_t dims[5] = {chq, (uint64_t)a->W, (uint64_t)a->V, (uint64_t)a->H, (uint64_t)a->F};
    const uint64_t ld = (uint64_t)a->q_ld;
    const uint64_t str[4] = {ld, ld * a->W, ld * a->W * a->V, ld * a->W * a->V * a->H};
    const uint32_t box[5] = {64u, (uint32_t)p.qw, 1u, (uint32_t)p.qh, 1u};
    int rc = cached_tmap_bf16(&p.mapQ, a->q, 5, dims, str, box, 128);
    if (rc != PN_OK) return rc;
  }
  {
    const uint64_t Fk = (uint64_t)((a->F + a->kv_frame_div - 1) / a->kv_frame_div);
    const uint64_t dims[5] = {chq, (uint64_t)a->Wk, (uint64_t)a->Vk, (uint64_t)a->Hk, Fk};
    const uint64_t ld = (uint64_t)a->kv_ld;
    const uint64_t str[4] = {ld, ld * a->Wk, ld * a->Wk * a->Vk, ld * a->Wk * a->Vk * a->Hk};
    const uint32_t box[5] = {64u, (uint32_t)p.kw, 1u, (uint32_t)p.kh, 1u};
    int rc = cached_tmap_bf16(&p.mapK, a->k, 5, dims, str, box, 128);
    if (rc != PN_OK) return rc;
    rc = cached_tmap_bf16(&p.mapV, a->v, 5, dims, str, box, 128);
    if (rc != PN_OK) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    for (int i = 0; i < 10; ++i)
      PN_CHECK_CUDA(cudaFuncSetAttribute(fa_kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_TOTAL));
    attr_set = true;
  }
  const long long tiles = (long long)p.tiles_x * p.tiles_y * a->heads * a->V * a->F;
  PN_REQUIRE(tiles > 0 && tiles < (1ll << 31), "pn_attention: too many query tiles");
  p.total_tiles = (int)tiles;
  const int grid = tiles < sm_count() ? (int)tiles : sm_count();
  const int nch = p.kv_n / 16;
  const int slot = nch == 8 ? 4 : nch == 7 ? 3 : nch == 5 ? 2 : nch == 2 ? 1 : 0;
  fa_kernels[(p.kv_rows < p.kv_n ? 5 : 0) + slot]<<<grid, FA_THREADS, FA_SMEM_TOTAL, reinterpret_cast<cudaStream_t>(stream_v)>>>(p);
  PN_CHECK_CUDA(cudaGetLastError());
  return PN_OK;
}
