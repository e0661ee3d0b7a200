"""Op layer: torch CUDA tensors in, hand-written sm_100a kernels (via the C ABI) out.

Every method validates shapes/dtypes, builds the plain-C argument struct and launches on torch's current
stream. No method computes anything in torch: torch is only the allocator and stream owner here.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

BF16 = torch.bfloat16
F32 = torch.float32


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(cond, msg):
    if not cond:
        raise ValueError(msg)


class NativeOps:
    """The production op set. `launches` counts kernel launches issued through the C ABI."""

    def __init__(self):
        self.lib = _lib.load()
        self.launches = 0

    # ------------------------------------------------------------------ GEMM / implicit conv
    def gemm(self, a, w, *, bias=None, rowvec=None, rows_per_group=0, n_groups=0, residual=None,
             geglu=False, out_dtype=F32, taps=(1, 1), out=None):
        """out[row, :] = epi(sum_taps A[shifted pixel] @ w^T). See include/panacea_b200.h::pn_gemm.

        a: bf16 [..., C] (taps == (1,1): any leading dims, rows may be strided views with C contiguous)
           or bf16 [NB, H, W, C] for taps (3,3) / (3,1).
        w: bf16 [N, taps_h*taps_w*C].
        """
        _req(a.is_cuda and a.dtype == BF16 and w.dtype == BF16, "gemm: a and w must be CUDA bf16")
        _req(a.stride(-1) == 1 and w.is_contiguous(), "gemm: innermost dim must be contiguous")
        th, tw = taps
        Cc = a.shape[-1]
        N = w.shape[0]
        _req(w.shape[1] == th * tw * Cc, f"gemm: weight K {w.shape[1]} != taps*C {th * tw * Cc}")
        if (th, tw) == (1, 1):
            lead = a.shape[:-1]
            a2 = a.reshape(-1, Cc) if a.is_contiguous() else a
            if a2.dim() != 2:
                # strided view: collapse leading dims only if they are uniformly strided
                rows = 1
                for s in lead:
                    rows *= s
                st = a.stride(-2)
                ok = all(a.stride(i) == a.stride(i + 1) * a.shape[i + 1] for i in range(a.dim() - 2))
                _req(ok, "gemm: cannot flatten strided A view")
                a2 = a.as_strided((rows, Cc), (st, 1))
            NB, H, W = 1, 1, a2.shape[0]
            sw = a2.stride(0)
            sh = sw * W
            sn = sh
            a_use = a2
        else:
            _req(a.dim() == 4, "gemm: conv mode expects [NB,H,W,C]")
            NB, H, W = a.shape[0], a.shape[1], a.shape[2]
            sn, sh, sw = a.stride(0), a.stride(1), a.stride(2)
            lead = a.shape[:-1]
            a_use = a
        rows = NB * H * W
        n_out = N // 2 if geglu else N
        if out is None:
            out = torch.empty((rows, n_out), device=a.device, dtype=out_dtype)
        else:
            _req(out.dtype == out_dtype and out.stride(-1) == 1, "gemm: bad out tensor")
        out2 = out.reshape(rows, -1) if out.is_contiguous() else out
        args = _lib.GemmArgs()
        args.A = a_use.data_ptr(); args.B = w.data_ptr(); args.out = out2.data_ptr()
        args.bias = None if bias is None else bias.data_ptr()
        args.rowvec = None if rowvec is None else rowvec.data_ptr()
        args.residual = None if residual is None else residual.data_ptr()
        if bias is not None:
            _req(bias.dtype == F32 and bias.numel() == N, "gemm: bias must be fp32 [N]")
        if rowvec is not None:
            _req(rowvec.dtype == F32 and rowvec.is_contiguous() and rowvec.shape[-1] == N, "gemm: rowvec fp32 [G,N]")
            _req(rows_per_group > 0 and n_groups > 0 and rowvec.numel() == n_groups * N, "gemm: rowvec groups")
        if residual is not None:
            _req(residual.dtype == F32 and residual.stride(-1) == 1, "gemm: residual must be fp32")
            r2 = residual.reshape(rows, -1) if residual.is_contiguous() else residual
            args.ldr = r2.stride(0)
        args.NB, args.H, args.W, args.C = NB, H, W, Cc
        args.a_stride_w, args.a_stride_h, args.a_stride_n = sw, sh, sn
        args.ldo = out2.stride(0)
        args.N = N; args.taps_h = th; args.taps_w = tw
        args.rows_per_group = rows_per_group; args.n_groups = n_groups
        args.out_bf16 = 1 if out_dtype == BF16 else 0
        args.geglu = 1 if geglu else 0
        _lib.check(self.lib.pn_gemm(C.byref(args), _stream()), "pn_gemm")
        self.launches += 1
        return out.reshape(*lead, n_out) if out.is_contiguous() else out
