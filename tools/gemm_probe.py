"""Timing probe for the tcgen05 GEMM under the experiment switches (PN_GEMM_DEBUG, PN_CONV_HALO, PN_GEMM_MODE)."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from panacea_b200.ops import NativeOps  # noqa: E402
from tools.bench_kernels import timeit  # noqa: E402

BF16 = torch.bfloat16


def main():
    ops = NativeOps()
    dev = "cuda"
    tag = " ".join(f"{k}={os.environ[k]}" for k in ("PN_GEMM_DEBUG", "PN_CONV_HALO", "PN_GEMM_MODE") if k in os.environ) or "default"
    x = torch.randn(16, 32, 336, 320, device=dev).to(BF16)
    wc = (torch.randn(320, 9 * 320, device=dev) * 0.02).to(BF16)
    t = timeit(lambda: ops.gemm(x, wc, taps=(3, 3)))
    print(f"[{tag}] conv3x3 L0 N=320 K=2880: {t*1e6:8.1f} us  {2.0*172032*2880*320/t/1e12:7.1f} TF/s")
    xa = torch.randn(172032, 2880, device=dev).to(BF16)
    t = timeit(lambda: ops.gemm(xa, wc))
    print(f"[{tag}] plain  M=172032 N=320 K=2880: {t*1e6:8.1f} us  {2.0*172032*2880*320/t/1e12:7.1f} TF/s")
    wd = (torch.randn(2560, 2880, device=dev) * 0.02).to(BF16)
    t = timeit(lambda: ops.gemm(xa[:43008], wd, out_dtype=BF16))
    print(f"[{tag}] plain  M=43008 N=2560 K=2880 (BN=256): {t*1e6:8.1f} us  {2.0*43008*2880*2560/t/1e12:7.1f} TF/s")
    del xa
    a = torch.randn(8192, 8192, device=dev).to(BF16)
    b = torch.randn(8192, 8192, device=dev).to(BF16)
    t = timeit(lambda: ops.gemm(a, b, out_dtype=BF16), iters=5)
    print(f"[{tag}] gemm 8192^3: {t*1e6:8.1f} us  {2.0*8192**3/t/1e12:7.1f} TF/s")


if __name__ == "__main__":
    main()
