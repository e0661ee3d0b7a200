"""Phase durations of the view-attention kernel's CTA 0 from its clock64 timeline (diagnostics build:
PN_GEMM_ROLE_TIMERS=1 python -m panacea_b200.build --force; run with PN_ATTN_DEBUG=8)."""
import ctypes, os, sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ.setdefault("PN_ATTN_DEBUG", "8")
from panacea_b200.ops import NativeOps
NEIGH = ((5, 1), (0, 2), (1, 3), (2, 4), (3, 5), (4,))
ops = NativeOps()
qkv = torch.randn(16, 32, 6, 56, 960, device="cuda").to(torch.bfloat16)
for cross in (False, True):
    for _ in range(2):
        ops.attention_view(qkv, 5, cross, NEIGH)
    buf = (ctypes.c_longlong * (3 * 96 * 8))()
    ops.lib.pn_debug_attn_timeline(buf, 3 * 96 * 8)
    t = np.array(buf, dtype=np.int64).reshape(3, 96, 8)
    lo, hi = 20, 80                                   # steady-state blocks of CTA 0
    names = ["wait S (from P stored of the previous block)", "S -> registers", "row maximum", "rendezvous", "exponentials", "wait previous PV", "store P"]
    print(f"--- {'cross' if cross else 'intra'}-view, CTA 0, blocks {lo}..{hi} (cycles, mean)")
    for g in (0, 1):
        a = t[g, lo:hi]
        prev_end = t[g, lo - 1:hi - 1, 6]
        d = [a[:, 0] - prev_end] + [a[:, e] - a[:, e - 1] for e in range(1, 7)]
        per = (a[1:, 6] - a[:-1, 6]).mean()
        print(f" group {'AB'[g]}: period {per:7.0f} | " + " | ".join(f"{n}: {x.mean():6.0f}" for n, x in zip(names, d)))
    m = t[2, lo:hi]
    print(f" issuer : S_A->S_B {np.mean(m[:,1]-m[:,0]):6.0f} | S_B->PV_A {np.mean(m[:,2]-m[:,1]):6.0f} | PV_A->PV_B {np.mean(m[:,3]-m[:,2]):6.0f} | PV_B->next S_A {np.mean(m[1:,0]-m[:-1,3]):6.0f}")
    print(f" offsets: group B 'S ready' minus group A 'S ready' {np.mean(t[1,lo:hi,0]-t[0,lo:hi,0]):6.0f}; exp start B - A {np.mean(t[1,lo:hi,3]-t[0,lo:hi,3]):6.0f}")
