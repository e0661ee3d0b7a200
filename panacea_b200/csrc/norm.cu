// Normalisation kernels (HBM-bound): they read the fp32 residual stream once and emit the bf16 MMA operand
// of the GEMM/conv that follows, with the affine transform and SiLU fused in.
//
//  * spatial GroupNorm(32)  — statistics over (C/32 channels x all H x Wtot pixels of one frame), i.e. over
//    ALL SIX VIEWS jointly (reference: util.py:276-283 eps 1e-5 in ResBlock3D, attention.py:129-132 eps 1e-6
//    in SpatialTemporalTransformer);
//  * pixel-wise GroupNorm(32) over (C/32 channels x T frames) per pixel (reference: openaimodel.py:509-515,
//    534-539: GroupNorm applied to the [(b h w), C, T] rearrangement) — here computed in place on the
//    [b, T, P, C] layout, no rearrange copies;
//  * LayerNorm(C) per token (attention.py:699-701, eps 1e-5).
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/panacea_b200.h"

namespace pn {

constexpr int GN_GROUPS = 32;
constexpr int GN_CHUNK = 64;  // pixels per partial-statistics block

// ---------------------------------------------------------------- spatial GN: partial sums per pixel chunk
// partial[f][chunk][g][2] = (sum, sumsq) over the chunk's pixels x group channels (fp32, <= 64*cpg terms each).
// Thread layout: (pixel lane, float4 column). With C = 320 a block of 256 threads runs 3 pixel lanes x 80 columns,
// so (almost) every thread streams; per-lane column sums meet in shared memory and are reduced in a fixed order
// (bit-reproducible, no atomics).
__global__ void __launch_bounds__(256) gn_partial_kernel(const float* __restrict__ x, float* __restrict__ partial,
                                                         int P, int C, int nchunks, int PL) {
  extern __shared__ float colacc[];   // [PL][2][C]
  const int f = blockIdx.y, chunk = blockIdx.x;
  const int cpg = C / GN_GROUPS;
  const int p0 = chunk * GN_CHUNK;
  const int p1 = min(P, p0 + GN_CHUNK);
  const int c4n = C / 4;
  const float* base = x + ((size_t)f * P) * C;
  const int cols = c4n < 256 ? c4n : 256;
  const int pl = threadIdx.x / cols;
  if (pl < PL) {
    for (int c4 = threadIdx.x - pl * cols; c4 < c4n; c4 += cols) {
      float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
      for (int p = p0 + pl; p < p1; p += PL) {
        const float4 v = *reinterpret_cast<const float4*>(base + (size_t)p * C + c4 * 4);
        s[0] += v.x; q[0] += v.x * v.x;
        s[1] += v.y; q[1] += v.y * v.y;
        s[2] += v.z; q[2] += v.z * v.z;
        s[3] += v.w; q[3] += v.w * v.w;
      }
      *reinterpret_cast<float4*>(colacc + (size_t)pl * 2 * C + c4 * 4) = make_float4(s[0], s[1], s[2], s[3]);
      *reinterpret_cast<float4*>(colacc + (size_t)pl * 2 * C + C + c4 * 4) = make_float4(q[0], q[1], q[2], q[3]);
    }
  }
  __syncthreads();
  if (threadIdx.x < GN_GROUPS * 2) {
    const int g = threadIdx.x >> 1, which = threadIdx.x & 1;
    float acc = 0.f;
    for (int l = 0; l < PL; ++l) {
      const float* src = colacc + (size_t)l * 2 * C + which * C + g * cpg;
      for (int j = 0; j < cpg; ++j) acc += src[j];
    }
    partial[((size_t)f * nchunks + chunk) * GN_GROUPS * 2 + threadIdx.x] = acc;
  }
}

// scale[f][c] = rstd*gamma[c]; shift[f][c] = beta[c] - mean*rstd*gamma[c]   (double-precision combine)
__global__ void gn_finalize_kernel(const float* __restrict__ partial, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ scale,
                                   float* __restrict__ shift, int P, int C, int nchunks, float eps) {
  __shared__ float s_mean[GN_GROUPS], s_rstd[GN_GROUPS];
  __shared__ double s_part[4][GN_GROUPS * 2];
  const int f = blockIdx.x;
  const int cpg = C / GN_GROUPS;
  {
    // 256 threads: (chunk quarter, group, sum|sumsq); double-precision combine in a fixed order
    const int slot = threadIdx.x & 63, quarter = threadIdx.x >> 6;
    double acc = 0.0;
    const float* pp = partial + (size_t)f * nchunks * GN_GROUPS * 2 + slot;
    for (int k = quarter; k < nchunks; k += 4) acc += (double)pp[(size_t)k * GN_GROUPS * 2];
    s_part[quarter][slot] = acc;
  }
  __syncthreads();
  if (threadIdx.x < GN_GROUPS) {
    const int g = threadIdx.x;
    const double s = (s_part[0][2 * g] + s_part[1][2 * g]) + (s_part[2][2 * g] + s_part[3][2 * g]);
    const double q = (s_part[0][2 * g + 1] + s_part[1][2 * g + 1]) + (s_part[2][2 * g + 1] + s_part[3][2 * g + 1]);
    const double n = (double)P * cpg;
    const double mean = s / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[threadIdx.x] = (float)mean;
    s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float sc = s_rstd[g] * gamma[c];
    scale[(size_t)f * C + c] = sc;
    shift[(size_t)f * C + c] = beta[c] - s_mean[g] * sc;
  }
}

// y = act(x*scale[f,c] + shift[f,c]) -> bf16 ; optional raw bf16 copy of x (input of the 1x1 skip conv).
// Thread layout: (pixel lane, 8-channel column) fixed per thread, so scale/shift live in registers and the loop has
// no integer division; a block covers GN_APPLY_PIX pixels of one frame.
constexpr int GN_APPLY_PIX = 64;
__global__ void __launch_bounds__(256) gn_apply_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                       const float* __restrict__ shift,
                                                       __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ raw,
                                                       int P, int C, int act_silu) {
  const int f = blockIdx.y;
  const int c8n = C / 8;
  const int cols = c8n < 256 ? c8n : 256;
  const int PL = 256 / cols;
  const int pl = threadIdx.x / cols;
  if (pl >= PL) return;
  const int p0 = blockIdx.x * GN_APPLY_PIX;
  const int p1 = min(P, p0 + GN_APPLY_PIX);
  const float* xb = x + (size_t)f * P * C;
  __nv_bfloat16* yb = y + (size_t)f * P * C;
  __nv_bfloat16* rb = raw ? raw + (size_t)f * P * C : nullptr;
  for (int c8 = threadIdx.x - pl * cols; c8 < c8n; c8 += cols) {
    float sc[8], sh[8];
    {
      const float4 a = *reinterpret_cast<const float4*>(scale + (size_t)f * C + c8 * 8);
      const float4 b = *reinterpret_cast<const float4*>(scale + (size_t)f * C + c8 * 8 + 4);
      const float4 c = *reinterpret_cast<const float4*>(shift + (size_t)f * C + c8 * 8);
      const float4 d = *reinterpret_cast<const float4*>(shift + (size_t)f * C + c8 * 8 + 4);
      sc[0] = a.x; sc[1] = a.y; sc[2] = a.z; sc[3] = a.w; sc[4] = b.x; sc[5] = b.y; sc[6] = b.z; sc[7] = b.w;
      sh[0] = c.x; sh[1] = c.y; sh[2] = c.z; sh[3] = c.w; sh[4] = d.x; sh[5] = d.y; sh[6] = d.z; sh[7] = d.w;
    }
    for (int p = p0 + pl; p < p1; p += PL) {
      const size_t off = (size_t)p * C + (size_t)c8 * 8;
      const float4 a = *reinterpret_cast<const float4*>(xb + off);
      const float4 b = *reinterpret_cast<const float4*>(xb + off + 4);
      float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      if (rb) {
        *reinterpret_cast<uint4*>(rb + off) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                          pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = v[j] * sc[j] + sh[j];
        v[j] = act_silu ? silu(t) : t;
      }
      *reinterpret_cast<uint4*>(yb + off) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                        pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
  }
}

// ---------------------------------------------------------------- pixel-wise temporal GN (+SiLU) -> bf16
// x: fp32 [b, T, P, C]. One block per (b, p); thread (t, g) = (warp, lane) owns the cpg channels of group g at frame t
// and keeps them in registers: a single pass over HBM, exact two-pass statistics (mean, then centred sum of squares)
// combined across the T warps through shared memory. Lanes read adjacent cpg-float runs, i.e. whole rows coalesced.
template <int CPG>
__global__ void __launch_bounds__(512) gn_pixel_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                                       int T, int P, int C, float eps, int act_silu) {
  __shared__ float red[16][32];
  const int t = threadIdx.x >> 5, g = threadIdx.x & 31;
  const int b = blockIdx.x / P, p = blockIdx.x - b * P;
  const size_t off = (((size_t)b * T + t) * P + p) * C + (size_t)g * CPG;
  float v[CPG];
  const float2* src = reinterpret_cast<const float2*>(x + off);
#pragma unroll
  for (int j = 0; j < CPG / 2; ++j) {
    const float2 a = src[j];
    v[2 * j] = a.x;
    v[2 * j + 1] = a.y;
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < CPG; ++j) s += v[j];
  red[t][g] = s;
  __syncthreads();
  float tot = 0.f;
  for (int k = 0; k < T; ++k) tot += red[k][g];
  const float n = (float)(T * CPG);
  const float mean = tot / n;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < CPG; ++j) {
    const float d = v[j] - mean;
    q += d * d;
  }
  __syncthreads();
  red[t][g] = q;
  __syncthreads();
  float qt = 0.f;
  for (int k = 0; k < T; ++k) qt += red[k][g];
  const float rstd = rsqrtf(qt / n + eps);
  uint32_t* dst = reinterpret_cast<uint32_t*>(y + off);
  const float2* g2 = reinterpret_cast<const float2*>(gamma + g * CPG);
  const float2* b2 = reinterpret_cast<const float2*>(beta + g * CPG);
#pragma unroll
  for (int j = 0; j < CPG / 2; ++j) {
    const float2 gm = g2[j], bt = b2[j];
    float a0 = (v[2 * j] - mean) * rstd * gm.x + bt.x;
    float a1 = (v[2 * j + 1] - mean) * rstd * gm.y + bt.y;
    if (act_silu) { a0 = silu(a0); a1 = silu(a1); }
    dst[j] = pack_bf16x2(a0, a1);
  }
}

// ---------------------------------------------------------------- LayerNorm per token -> bf16
// one warp per row; the row (C <= 2048 floats) lives in registers between the two passes.
template <int MAXV>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                                        long long rows, int C, float eps) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * C);
  const int n4 = C / 4;
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int j = lane + i * 32;
    if (j < n4) {
      v[i] = xr[j];
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int j = lane + i * 32;
    if (j < n4) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)C + eps);
  uint2* yr = reinterpret_cast<uint2*>(y + row * C);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int j = lane + i * 32;
    if (j < n4) {
      const float4 g = g4[j], bb = b4[j];
      const float o0 = (v[i].x - mean) * rstd * g.x + bb.x;
      const float o1 = (v[i].y - mean) * rstd * g.y + bb.y;
      const float o2 = (v[i].z - mean) * rstd * g.z + bb.z;
      const float o3 = (v[i].w - mean) * rstd * g.w + bb.w;
      yr[j] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
    }
  }
}

}  // namespace pn

using namespace pn;

extern "C" int64_t pn_groupnorm_workspace_floats(int64_t frames, int64_t pixels, int64_t channels) {
  const int64_t nchunks = (pixels + GN_CHUNK - 1) / GN_CHUNK;
  return frames * nchunks * GN_GROUPS * 2 + 2 * frames * channels;
}

extern "C" int pn_groupnorm_silu(const float* x, const float* gamma, const float* beta, void* y_bf16,
                                 void* raw_bf16, float* workspace, int64_t frames, int64_t pixels,
                                 int64_t channels, float eps, int act_silu, void* stream_v) {
  PN_REQUIRE(x && gamma && beta && y_bf16 && workspace, "pn_groupnorm_silu: null pointer");
  PN_REQUIRE(channels % 32 == 0 && channels % 8 == 0 && channels <= 8192, "pn_groupnorm_silu: C=%lld unsupported",
             (long long)channels);
  PN_REQUIRE(frames > 0 && pixels > 0, "pn_groupnorm_silu: empty input");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
  const int P = (int)pixels, C = (int)channels, F = (int)frames;
  const int nchunks = (P + GN_CHUNK - 1) / GN_CHUNK;
  float* partial = workspace;
  float* scale = workspace + (size_t)F * nchunks * GN_GROUPS * 2;
  float* shift = scale + (size_t)F * C;
  const int c4n = C / 4;
  const int PL = c4n < 256 ? 256 / c4n : 1;
  gn_partial_kernel<<<dim3(nchunks, F), 256, (size_t)PL * 2 * C * sizeof(float), st>>>(x, partial, P, C, nchunks, PL);
  PN_CHECK_CUDA(cudaGetLastError());
  gn_finalize_kernel<<<F, 256, 0, st>>>(partial, gamma, beta, scale, shift, P, C, nchunks, eps);
  PN_CHECK_CUDA(cudaGetLastError());
  const int gx = (P + GN_APPLY_PIX - 1) / GN_APPLY_PIX;
  gn_apply_kernel<<<dim3(gx, F), 256, 0, st>>>(x, scale, shift, reinterpret_cast<__nv_bfloat16*>(y_bf16),
                                                reinterpret_cast<__nv_bfloat16*>(raw_bf16), P, C, act_silu);
  PN_CHECK_CUDA(cudaGetLastError());
  return PN_OK;
}

extern "C" int pn_groupnorm_pixel_silu(const float* x, const float* gamma, const float* beta, void* y_bf16,
                                       int64_t batch, int64_t frames_per_seq, int64_t pixels, int64_t channels,
                                       float eps, int act_silu, void* stream_v) {
  PN_REQUIRE(x && gamma && beta && y_bf16, "pn_groupnorm_pixel_silu: null pointer");
  PN_REQUIRE(channels % 64 == 0, "pn_groupnorm_pixel_silu: C=%lld must be a multiple of 64", (long long)channels);
  PN_REQUIRE(batch > 0 && frames_per_seq > 0 && pixels > 0, "pn_groupnorm_pixel_silu: empty input");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
  PN_REQUIRE(frames_per_seq <= 16, "pn_groupnorm_pixel_silu: T=%lld > 16 unsupported", (long long)frames_per_seq);
  const long long blocks = batch * pixels;
  PN_REQUIRE(blocks < (1ll << 31), "pn_groupnorm_pixel_silu: grid too large");
  const int threads = (int)frames_per_seq * 32;
  __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(y_bf16);
  const int T = (int)frames_per_seq, P = (int)pixels, C = (int)channels;
  switch (C / 32) {
#define PN_GNP_CASE(CPG) case CPG: gn_pixel_kernel<CPG><<<(unsigned)blocks, threads, 0, st>>>(x, gamma, beta, y, T, P, C, eps, act_silu); break;
    PN_GNP_CASE(2) PN_GNP_CASE(4) PN_GNP_CASE(6) PN_GNP_CASE(8) PN_GNP_CASE(10) PN_GNP_CASE(12) PN_GNP_CASE(16) PN_GNP_CASE(20)
    PN_GNP_CASE(24) PN_GNP_CASE(30) PN_GNP_CASE(32) PN_GNP_CASE(40) PN_GNP_CASE(60) PN_GNP_CASE(80)
#undef PN_GNP_CASE
    default:
      return fail(PN_ERR_UNSUPPORTED, "pn_groupnorm_pixel_silu: C=%lld (C/32=%lld channels per group) is not instantiated",
                  (long long)channels, (long long)(channels / 32));
  }
  PN_CHECK_CUDA(cudaGetLastError());
  return PN_OK;
}

extern "C" int pn_layernorm(const float* x, const float* gamma, const float* beta, void* y_bf16, int64_t rows,
                            int64_t channels, float eps, void* stream_v) {
  PN_REQUIRE(x && gamma && beta && y_bf16, "pn_layernorm: null pointer");
  PN_REQUIRE(channels % 4 == 0 && channels <= 2048 && channels > 0, "pn_layernorm: C=%lld unsupported", (long long)channels);
  PN_REQUIRE(rows > 0, "pn_layernorm: empty input");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
  const long long blocks = (rows * 32 + 255) / 256;
  __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(y_bf16);
  const int C = (int)channels;
  if (C <= 512) layernorm_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(x, gamma, beta, y, rows, C, eps);
  else if (C <= 1024) layernorm_kernel<8><<<(unsigned)blocks, 256, 0, st>>>(x, gamma, beta, y, rows, C, eps);
  else layernorm_kernel<16><<<(unsigned)blocks, 256, 0, st>>>(x, gamma, beta, y, rows, C, eps);
  PN_CHECK_CUDA(cudaGetLastError());
  return PN_OK;
}
