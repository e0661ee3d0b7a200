"""Epilogue/mainloop decomposition of the K=320 transformer GEMMs at the L0 shape (PN_GEMM_DEBUG = 0..4), and the
per-role cycle accounting of CTA 0 (PN_GEMM_DEBUG=5). All PN_GEMM_DEBUG / PN_ATTN_DEBUG switches only exist in a
diagnostics build: PN_GEMM_ROLE_TIMERS=1 python -m panacea_b200.build --force (the product build folds them away)."""
import ctypes, os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from panacea_b200.ops import NativeOps
from tools.bench_kernels import timeit
ops = NativeOps()
tag = " ".join(f"{k}={os.environ[k]}" for k in ("PN_GEMM_DEBUG", "PN_GEMM_BN") if k in os.environ)
M = 172032
a = torch.randn(M, 320, device="cuda").to(torch.bfloat16)
a4 = torch.randn(M, 1280, device="cuda").to(torch.bfloat16)
res = torch.randn(M, 320, device="cuda")
res16 = res.to(torch.bfloat16)
def w(n, k): return (torch.randn(n, k, device="cuda") * 0.02).to(torch.bfloat16)
w_lin, w_qkv, w_ff1, w_ff2 = w(320, 320), w(960, 320), w(2560, 320), w(320, 1280)
b320 = torch.randn(320, device="cuda"); b2560 = torch.randn(2560, device="cuda")
rows = [
    ("linear+res fp32", lambda: ops.gemm(a, w_lin, bias=b320, residual=res), 2.0 * M * 320 * 320),
    ("qkv bf16", lambda: ops.gemm(a, w_qkv, out_dtype=torch.bfloat16), 2.0 * M * 960 * 320),
    ("ff1 geglu", lambda: ops.gemm(a, w_ff1, bias=b2560, geglu=True, out_dtype=torch.bfloat16), 2.0 * M * 2560 * 320),
    ("ff2+res fp32", lambda: ops.gemm(a4, w_ff2, bias=b320, residual=res), 2.0 * M * 320 * 1280),
    ("to_out+res bf16", lambda: ops.gemm(a, w_lin, bias=b320, residual=res16, out_dtype=torch.bfloat16), 2.0 * M * 320 * 320),
    ("ff2+res bf16", lambda: ops.gemm(a4, w_ff2, bias=b320, residual=res16, out_dtype=torch.bfloat16), 2.0 * M * 320 * 1280),
]
for name, fn, fl in rows:
    t = timeit(fn)
    print(f"[{tag}] {name:18s} {t*1e6:8.1f} us {fl/t/1e12:7.1f} TF/s")
    if int(os.environ.get("PN_GEMM_DEBUG", "0")) in (5, 8, 9, 10, 11, 12, 14, 15):
        c = (ctypes.c_ulonglong * 16)()
        ops.lib.pn_debug_gemm_counters(c)
        v = [int(x) for x in c]
        nt = max(v[3], 1)
        print(f"      per tile (CTA 0, {v[3]} tiles): issuer {v[0]/nt:.0f} cyc (wait accumulator {v[1]/nt:.0f}, wait operands {v[2]/nt:.0f}); "
              f"producer {v[4]/nt:.0f} (wait slots {v[5]/nt:.0f}); epilogue warp {v[6]/nt:.0f} (wait tmem_full {v[7]/nt:.0f}, "
              f"wait chunk {v[8]/nt:.0f}); store warp waits: chunks {v[9]/nt:.0f}, smem reads {v[10]/nt:.0f}")
