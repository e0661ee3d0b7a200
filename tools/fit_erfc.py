"""Coefficients of the one-MUFU erfc used by the GEGLU epilogue (panacea_b200/csrc/ptx.cuh::geglu_f32x2):
Phi(-t) = 0.5 * 2^(-t Q(t)), Q of degree 4 fitted on [0, 7] with weights that minimise the ABSOLUTE error of Phi."""
import numpy as np
from scipy.special import erfc

t = np.linspace(1e-6, 7.0, 200001)
q = 0.5 * erfc(t / np.sqrt(2))
y = -np.log2(2 * q) / t
deg = 4
A = np.vander(t, deg + 1, increasing=True)
w, wt = q * t, np.ones_like(t)
for _ in range(60):                       # Lawson iteration towards the minimax fit
    c = np.linalg.lstsq(A * (w * wt)[:, None], y * w * wt, rcond=None)[0]
    err = np.abs(0.5 * np.exp2(-t * (A @ c)) - q)
    wt = wt * (0.5 + err / err.max())
    wt /= wt.max()
t32, c32 = t.astype(np.float32), c.astype(np.float32)
acc = np.full_like(t32, c32[-1])
for k in range(deg - 1, -1, -1):
    acc = (acc * t32 + c32[k]).astype(np.float32)
q32 = (np.float32(0.5) * np.exp2((-t32 * acc).astype(np.float32))).astype(np.float32)
print("coefficients c0..c4:", [float(x) for x in c32])
tt = np.linspace(7.0, 1e4, 100001)
print("min of t Q(t) beyond the fitted range (must stay large: the kernel does not clamp |g|):", float((np.polyval(c[::-1], tt) * tt).min()))
print(f"max |Phi error| exact arithmetic {err.max():.2e}, fp32 Horner {np.abs(q32 - q).max():.2e}, "
      f"max |gelu error| {np.abs(t * (q32 - q)).max():.2e}")
