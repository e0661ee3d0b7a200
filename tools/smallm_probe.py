"""Level-3 / mid-block 3x3 convs (rows = 16 x 4 x 42 = 2688): single-CTA 128 x 160 tiles (default for < 2 waves of tiles)
vs CTA pairs (PN_GEMM_MODE=2)."""
import os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from panacea_b200.ops import NativeOps
from tools.bench_kernels import timeit
ops = NativeOps()
for name, Fr, H, W, C, N in (("L3 conv 1280->1280", 16, 4, 42, 1280, 1280), ("L3 conv 2560->1280", 16, 4, 42, 2560, 1280), ("L2 conv 2560->1280", 16, 8, 84, 2560, 1280),
                             ("L3 linear 1280->1280", 1, 1, 2688, 1280, 1280)):
    x = torch.randn(Fr, H, W, C, device="cuda").to(torch.bfloat16)
    taps = (1, 1) if "linear" in name else (3, 3)
    w = (torch.randn(N, taps[0] * taps[1] * C, device="cuda") * 0.01).to(torch.bfloat16)
    b = torch.randn(N, device="cuda")
    t = timeit(lambda: ops.gemm(x, w, bias=b, taps=taps), iters=20)
    fl = 2.0 * Fr * H * W * N * w.shape[1]
    print(f"[PN_GEMM_MODE={os.environ.get('PN_GEMM_MODE', '0')}] {name:22s} {t*1e6:7.1f} us  {fl/t/1e12:7.1f} TF/s", flush=True)
