"""(PN_GEMM_DEBUG switches need a diagnostics build: PN_GEMM_ROLE_TIMERS=1 python -m panacea_b200.build --force)
MMA-rate probe: plain GEMM M=86016, K=2880, N=3840 with a forced N tile (PN_GEMM_BN) in CTA-pair mode."""
import os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from panacea_b200.ops import NativeOps
from tools.bench_kernels import timeit
ops = NativeOps()
tag = " ".join(f"{k}={os.environ[k]}" for k in ("PN_GEMM_DEBUG", "PN_GEMM_BN") if k in os.environ)
a = torch.randn(86016, 2880, device="cuda").to(torch.bfloat16)
w = (torch.randn(3840, 2880, device="cuda") * 0.02).to(torch.bfloat16)
t = timeit(lambda: ops.gemm(a, w, out_dtype=torch.bfloat16))
print(f"[{tag}] M=86016 N=3840 K=2880: {t*1e6:8.1f} us {2.0*86016*3840*2880/t/1e12:7.1f} TF/s")
