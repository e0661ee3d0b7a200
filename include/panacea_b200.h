/*
 * panacea_b200 — C ABI of the B200-native Panacea denoising hot path.
 *
 * Drop-in boundary. The reference is pure Python with no FFI of its own; the
 * interface this library stands behind is the `nn.Module.forward` surface of
 *   sgm/modules/diffusionmodules/wrappers.py:37-70     OpenAIWrapperControlLDM3D.forward
 *   sgm/modules/diffusionmodules/controlmodel.py:86-202 ControlNet3D / ControlledUNetModel3D.forward
 *   sgm/modules/diffusionmodules/openaimodel.py:499-542 ResBlock3D._forward
 *   sgm/modules/attention.py:407-610,229-291,1064-1134  view / text / temporal attention, STT
 *   sgm/modules/diffusionmodules/sampling.py:96-133     EulerEDMSampler step
 * and every entry point below names the reference call site it replaces. The Python host
 * (panacea_b200/sgm/...) mirrors those classes and binds these symbols with ctypes.
 *
 * Conventions
 *  - plain pointers + sizes only; all tensor pointers are DEVICE pointers owned by the caller;
 *  - activations are channels-last: fp32 residual stream [frames, H, Wtot, C], bf16 MMA operands;
 *    frame index = b*T + t (t fastest), Wtot = 6 views side by side (view-major along W);
 *  - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it and allocation-free
 *    (CUDA-graph capturable); a handle-free design: no hidden device state except memoised TMA maps;
 *  - return 0 on success, negative pn_status otherwise; message via pn_last_error() (thread-local);
 *  - there is NO CPU fallback anywhere: without a CUDA device every compute entry point fails.
 */
#ifndef PANACEA_B200_H
#define PANACEA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PN_ABI_VERSION 1

enum pn_status { PN_STATUS_OK = 0, PN_STATUS_INVALID = -1, PN_STATUS_CUDA = -2, PN_STATUS_UNSUPPORTED = -3 };

const char* pn_last_error(void);
int pn_abi_version(void);

/* ------------------------------------------------------------------------------------------------
 * pn_gemm — tcgen05 GEMM / implicit-GEMM convolution (sm_100a, TMA + TMEM).
 * Replaces: nn.Linear (attention.py:94,113,220-226; openaimodel.py:939-941), nn.Conv2d 3x3 stride 1 over
 * the 6-view panorama (openaimodel.py:413,455-462,125), nn.Conv1d k=3 over frames (openaimodel.py:418,
 * 468-476), 1x1 skip / zero convs (openaimodel.py:486; controlmodel.py:81-84).
 *   out[row, n] = epi( sum_{th,tw,c} A[nb, y+th-taps_h/2, x+tw-taps_w/2, c] * B[n, (th*taps_w+tw)*C + c] )
 * with zero padding outside [0,H)x[0,W) and row = (nb*H + y)*W + x.
 * epi: + bias[n] + rowvec[(row / rows_per_group) % n_groups, n]; GEGLU (interleaved value/gate columns ->
 * N/2 outputs, attention.py:91-99); + residual[row, n] (fp32); store fp32 or bf16.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pn_gemm_args {
  const void* A;          /* bf16 [NB, H, W, C] with element strides below (C contiguous) */
  const void* B;          /* bf16 [N, taps_h*taps_w*C], K contiguous */
  void* out;              /* fp32 or bf16 [NB*H*W, ldo] */
  const float* bias;      /* [N] or NULL */
  const float* rowvec;    /* [n_groups, N] or NULL */
  const float* residual;  /* fp32 [NB*H*W, ldr] or NULL (may alias out when both fp32) */
  int64_t NB, H, W, C;
  int64_t a_stride_w, a_stride_h, a_stride_n; /* elements */
  int64_t ldo, ldr;
  int32_t N;
  int32_t taps_h, taps_w;
  int32_t rows_per_group, n_groups;
  int32_t out_bf16;
  int32_t geglu;
} pn_gemm_args;

int pn_gemm(const pn_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PANACEA_B200_H */
