"""(needs a diagnostics build: PN_GEMM_ROLE_TIMERS=1 python -m panacea_b200.build --force)
UMMA issue-rate probe: MMA-only GEMM (PN_GEMM_DEBUG=1) at several N tiles, 1-CTA (PN_GEMM_MODE=1) or CTA pairs (=2).
Prints cycles per tcgen05.mma (K=16) assuming the SM clock given by nvidia-smi at run time."""
import ctypes, os, subprocess, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from panacea_b200.ops import NativeOps
from tools.bench_kernels import timeit
ops = NativeOps()
tag = " ".join(f"{k}={os.environ[k]}" for k in ("PN_GEMM_DEBUG", "PN_GEMM_MODE", "PN_GEMM_BN") if k in os.environ)
ncta = 2 if os.environ.get("PN_GEMM_MODE") == "2" or "PN_GEMM_BN" in os.environ else 1
M, K = 86016, 2880
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
for N, bn in ((32, 32), (64, 64), (4096, 128), (3840, 160)):
    if "PN_GEMM_BN" in os.environ:
        bn = int(os.environ["PN_GEMM_BN"])
    w = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
    t = timeit(lambda: ops.gemm(a, w, out_dtype=torch.bfloat16))
    mhz = float(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.split()[0])
    units = 148 // ncta
    tiles = (M // (128 * ncta)) * ((N + bn - 1) // bn)
    mmas_per_unit = -(-tiles // units) * (K // 64) * 4
    print(f"[{tag}] N={N:5d} BN={bn:3d} ncta={ncta}: {t*1e6:8.1f} us  {2.0*M*N*K/t/1e12:7.1f} TF/s  ~{t*mhz*1e6/mmas_per_unit:6.1f} cycles/MMA @ {mhz:.0f} MHz (idle clock reading)")
