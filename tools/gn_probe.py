"""GroupNorm(+SiLU) at the level shapes of configs[1]: fused cooperative kernel vs its two phases as plain launches
(PN_GN_TWO_PHASE=1 under ncu shows the phases separately)."""
import os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from panacea_b200.ops import NativeOps
from tools.bench_kernels import timeit
ops = NativeOps()
for name, P, C in (("level 0", 32 * 336, 320), ("level 0 (concat 640)", 32 * 336, 640), ("level 1", 16 * 168, 640), ("level 2", 8 * 84, 1280), ("mid", 4 * 42, 1280)):
    x = torch.randn(16, P, C, device="cuda")
    g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
    silu = os.environ.get("PN_PROBE_NOSILU") != "1"
    t = timeit(lambda: ops.groupnorm(x, g, b, 1e-5, silu), iters=20)
    mb = x.numel() * 6 / 1e6
    sys.stdout.flush()
    print(f"[{'two-phase' if os.environ.get('PN_GN_TWO_PHASE') == '1' else 'fused'}] {name:22s} {t*1e6:7.1f} us  {mb/t/1e6:6.2f} TB/s (x fp32 once + y bf16)")
