// Cross-frame (temporal) self-attention: every pixel attends over its T <= 16 frames (attention.py:1116-1125 feeding
// CrossAttention.forward :229-291 with context=None). Sequences this short cannot fill a 128-row UMMA tile, and the
// op is HBM-bound (it reads q,k,v once and writes o once), so this is a CUDA-core kernel: one warp per
// (sequence b, pixel p, head); q/k/v rows are staged in shared memory, scores and softmax in fp32.
//
// Layout: qkv bf16 [b, T, P, ld] with q/k/v at channel offsets given by the three base pointers; the
// "(b t)(h w) c -> (b h w) t c" rearrangement of the reference is just this indexing — nothing is copied.
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/panacea_b200.h"

namespace pn {

constexpr int TA_MAXT = 16;
constexpr int TA_WARPS = 4;
constexpr int TA_PITCH = 72;  // 144 B rows: 16-B aligned and bank-shifted by 4 words per row (conflict-free)

__global__ void __launch_bounds__(TA_WARPS * 32) attn_temporal_kernel(const __nv_bfloat16* __restrict__ q,
                                                                      const __nv_bfloat16* __restrict__ k,
                                                                      const __nv_bfloat16* __restrict__ v,
                                                                      __nv_bfloat16* __restrict__ out, int nb, int T, int P,
                                                                      int heads, long long ld, long long out_ld, float scale) {
  __shared__ __align__(16) __nv_bfloat16 sq[TA_WARPS][TA_MAXT][TA_PITCH];
  __shared__ __align__(16) __nv_bfloat16 sk[TA_WARPS][TA_MAXT][TA_PITCH];
  __shared__ __align__(16) __nv_bfloat16 sv[TA_WARPS][TA_MAXT][TA_PITCH];
  __shared__ float sp[TA_WARPS][TA_MAXT][TA_MAXT + 1];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * TA_WARPS + w;
  const long long total = (long long)nb * P * heads;
  if (item >= total) return;   // whole warp exits together
  const int head = (int)(item % heads);
  const long long bp = item / heads;
  const int pix = (int)(bp % P);
  const int b = (int)(bp / P);
  // stage q,k,v rows: each row is 64 bf16 = 128 B = 8 x 16 B; lanes 0..7 -> row t, lanes 8..15 -> row t+1, ...
  for (int i = lane; i < T * 8; i += 32) {
    const int t = i >> 3, ch = i & 7;
    const long long tok = ((long long)(b * T + t) * P + pix);
    const long long off = tok * ld + head * 64 + ch * 8;
    *reinterpret_cast<uint4*>(&sq[w][t][ch * 8]) = *reinterpret_cast<const uint4*>(q + off);
    *reinterpret_cast<uint4*>(&sk[w][t][ch * 8]) = *reinterpret_cast<const uint4*>(k + off);
    *reinterpret_cast<uint4*>(&sv[w][t][ch * 8]) = *reinterpret_cast<const uint4*>(v + off);
  }
  __syncwarp();
  // scores: pair (i, j) per lane-iteration
  for (int pr = lane; pr < T * T; pr += 32) {
    const int i = pr / T, j = pr - i * T;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < 64; d += 2) {
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&sq[w][i][d]));
      const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&sk[w][j][d]));
      s += a.x * c.x + a.y * c.y;
    }
    sp[w][i][j] = s * scale;
  }
  __syncwarp();
  // softmax per row (T rows, lane < T handles one row)
  if (lane < T) {
    float mx = -INFINITY;
    for (int j = 0; j < T; ++j) mx = fmaxf(mx, sp[w][lane][j]);
    float sum = 0.f;
    for (int j = 0; j < T; ++j) {
      const float e = __expf(sp[w][lane][j] - mx);
      sp[w][lane][j] = e;
      sum += e;
    }
    const float inv = 1.f / sum;
    for (int j = 0; j < T; ++j) sp[w][lane][j] *= inv;
  }
  __syncwarp();
  // output: lane owns channels (2*lane, 2*lane+1)
  for (int i = 0; i < T; ++i) {
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < T; ++j) {
      const float pj = sp[w][i][j];
      const float2 vv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&sv[w][j][2 * lane]));
      o0 += pj * vv.x;
      o1 += pj * vv.y;
    }
    const long long tok = ((long long)(b * T + i) * P + pix);
    *reinterpret_cast<uint32_t*>(out + tok * out_ld + head * 64 + 2 * lane) = pack_bf16x2(o0, o1);
  }
}

}  // namespace pn

using namespace pn;

extern "C" int pn_attention_temporal(const void* q, const void* k, const void* v, void* out, int64_t batch, int64_t T,
                                     int64_t pixels, int32_t heads, int32_t head_dim, int64_t ld, int64_t out_ld,
                                     float scale, void* stream_v) {
  PN_REQUIRE(q && k && v && out, "pn_attention_temporal: null pointer");
  PN_REQUIRE(head_dim == 64, "pn_attention_temporal: head_dim %d unsupported (64 only)", head_dim);
  PN_REQUIRE(T >= 1 && T <= TA_MAXT, "pn_attention_temporal: T=%lld out of range 1..16", (long long)T);
  PN_REQUIRE(batch > 0 && pixels > 0 && heads > 0 && ld % 8 == 0 && out_ld % 2 == 0, "pn_attention_temporal: bad arguments");
  const long long total = batch * pixels * heads;
  const long long blocks = (total + TA_WARPS - 1) / TA_WARPS;
  PN_REQUIRE(blocks < (1ll << 31), "pn_attention_temporal: grid too large");
  attn_temporal_kernel<<<(unsigned)blocks, TA_WARPS * 32, 0, reinterpret_cast<cudaStream_t>(stream_v)>>>(
      reinterpret_cast<const __nv_bfloat16*>(q), reinterpret_cast<const __nv_bfloat16*>(k),
      reinterpret_cast<const __nv_bfloat16*>(v), reinterpret_cast<__nv_bfloat16*>(out), (int)batch, (int)T, (int)pixels, heads,
      ld, out_ld, scale);
  PN_CHECK_CUDA(cudaGetLastError());
  return PN_OK;
}
