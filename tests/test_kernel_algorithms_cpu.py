"""CPU restatements of two numerical schemes the CUDA kernels rely on, checked against plain fp64 math:
the lazy-rescale online softmax of attn_fa.cu and the frame-wave GroupNorm combine of norm.cu. They pin the ALGORITHMS
(what may be reordered / left stale without changing the result); the kernels themselves are checked on the GPU."""
import numpy as np
import torch


def _bf16(x):
    return torch.from_numpy(np.asarray(x, dtype=np.float32)).to(torch.bfloat16).float().numpy()


def test_lazy_rescale_online_softmax_is_exact_up_to_bf16_rounding_of_p():
    """attn_fa.cu: per key block p = 2^(s*c - m) with a reference maximum m that is only raised when the block maximum
    exceeds it by more than 2^8; l and O accumulate against the same (possibly stale) m, O is rescaled only then.
    Mathematically exact for any schedule of m; numerically it only adds the bf16 rounding of p (<= 2^-9 relative)."""
    rng = np.random.default_rng(0)
    Nq, Nk, D, blk, lazy = 64, 1792, 64, 112, 8.0
    q = _bf16(rng.standard_normal((Nq, D)))
    k = _bf16(rng.standard_normal((Nk, D)) * 1.5)        # spread the logits so that the maximum moves across blocks
    v = _bf16(rng.standard_normal((Nk, D)))
    c = (D ** -0.5) * 1.4426950408889634
    s_all = q.astype(np.float64) @ k.astype(np.float64).T
    m = np.full(Nq, -np.inf)
    l = np.zeros(Nq)
    o = np.zeros((Nq, D))
    raised = 0
    for j0 in range(0, Nk, blk):
        s = s_all[:, j0:j0 + blk]
        m_new = s.max(axis=1) * c
        upd = m_new > m + lazy                            # first block: m = -inf
        alpha = np.where(upd, np.exp2(np.where(np.isinf(m), -np.inf, m - m_new)), 1.0)
        m = np.where(upd, m_new, m)
        raised += int(upd.sum())
        p = np.exp2(s * c - m[:, None])
        assert p.max() <= 2.0 ** lazy * 1.0001            # stale maximum: values up to 2^8, never overflow
        l = l * alpha + p.sum(axis=1)
        o = o * alpha[:, None] + _bf16(p).astype(np.float64) @ v[j0:j0 + blk].astype(np.float64)
    out = o / l[:, None]
    w = np.exp(s_all * (D ** -0.5) - (s_all * (D ** -0.5)).max(axis=1, keepdims=True))
    ref = (w / w.sum(axis=1, keepdims=True)) @ v.astype(np.float64)
    assert raised < Nq * (Nk // blk) // 4                 # the rescale really is rare
    assert np.abs(out - ref).max() < 4e-3 * np.abs(ref).max() + 1e-3


def test_groupnorm_partial_combine_matches_two_pass_statistics():
    """norm.cu::gn_fused_kernel: every CTA of a frame sums (x, x^2) over its pixel range in fp32, the partials are
    combined in fp64 in a fixed order, var = E[x^2] - mean^2 clamped at 0. Against torch.group_norm in fp64."""
    g = torch.Generator().manual_seed(1)
    P, C, cpf, eps = 1000, 64, 7, 1e-5
    x = torch.randn(P, C, generator=g) * 3 + 1.5          # non-zero mean: the E[x^2]-mean^2 form must still hold up
    cpg = C // 32
    ppc = -(-P // cpf)
    part = torch.zeros(cpf, 32, 2, dtype=torch.float32)
    for r in range(cpf):
        xs = x[r * ppc:(r + 1) * ppc].float().reshape(-1, 32, cpg)
        part[r, :, 0] = xs.sum(dim=(0, 2))
        part[r, :, 1] = (xs * xs).sum(dim=(0, 2))
    tot = part.double().sum(0)
    n = P * cpg
    mean = tot[:, 0] / n
    var = (tot[:, 1] / n - mean * mean).clamp_min(0)
    y = (x.double().reshape(P, 32, cpg) - mean[None, :, None]) / torch.sqrt(var + eps)[None, :, None]
    ref = torch.nn.functional.group_norm(x.double().t().reshape(1, C, P), 32, eps=eps).reshape(C, P).t()
    torch.testing.assert_close(y.reshape(P, C), ref, rtol=1e-5, atol=1e-5)
