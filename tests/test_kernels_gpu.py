"""GPU parity of every non-GEMM kernel against plain torch fp32 math on the same inputs (same box, same GPU)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# the torch references below must be true fp32 (cuDNN/cuBLAS default to TF32 for conv/matmul on this GPU)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

NEIGH = ((5, 1), (0, 2), (1, 3), (2, 4), (3, 5), (4,))


@pytest.fixture(scope="module")
def ops():
    from panacea_b200.ops import NativeOps
    return NativeOps()


def _rand(shape, seed, scale=1.0, dtype=torch.float32, shift=0.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale + shift).to(dtype).cuda()


def _close(got, ref, tol, name):
    got = got.float()
    assert torch.isfinite(got).all(), f"{name}: non-finite"
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err <= tol * scale, f"{name}: max err {err:.4e} (scale {scale:.3e})"


# ------------------------------------------------------------------------------------------------ norms
@pytest.mark.parametrize("Fr,P,C", [(2, 100, 64), (3, 10752, 320), (16, 168, 1280), (2, 768, 1920), (1, 28, 2560)])
@pytest.mark.parametrize("silu", [False, True])
def test_groupnorm(ops, Fr, P, C, silu):
    x = _rand((Fr, P, C), 1, 2.0, shift=0.5)
    g = _rand((C,), 2, 0.1, shift=1.0); b = _rand((C,), 3, 0.1)
    eps = 1e-5 if silu else 1e-6
    y, raw = ops.groupnorm(x, g, b, eps, silu, want_raw=True)
    torch.cuda.synchronize()
    ref = F.group_norm(x.permute(0, 2, 1), 32, g, b, eps).permute(0, 2, 1)
    if silu:
        ref = F.silu(ref)
    _close(y, ref, 1e-2, "groupnorm")
    _close(raw, x, 1e-2, "groupnorm raw")


@pytest.mark.parametrize("b,T,P,C", [(2, 8, 50, 64), (1, 8, 2688, 320), (2, 4, 21, 1280), (1, 1, 64, 128)])
def test_groupnorm_pixel(ops, b, T, P, C):
    x = _rand((b, T, P, C), 4, 1.5, shift=-0.3)
    g = _rand((C,), 5, 0.1, shift=1.0); bb = _rand((C,), 6, 0.1)
    y = ops.groupnorm_pixel(x, g, bb, 1e-5, True)
    torch.cuda.synchronize()
    z = x.permute(0, 2, 3, 1).reshape(b * P, C, T)
    ref = F.silu(F.group_norm(z, 32, g, bb, 1e-5)).reshape(b, P, C, T).permute(0, 3, 1, 2)
    _close(y, ref, 1e-2, "groupnorm_pixel")


@pytest.mark.parametrize("rows,C", [(100, 64), (5000, 320), (777, 640), (300, 1280), (33, 128)])
def test_layernorm(ops, rows, C):
    x = _rand((rows, C), 7, 3.0, shift=1.0)
    g = _rand((C,), 8, 0.1, shift=1.0); b = _rand((C,), 9, 0.1)
    y = ops.layernorm(x, g, b, 1e-5)
    torch.cuda.synchronize()
    _close(y, F.layer_norm(x, (C,), g, b, 1e-5), 1e-2, "layernorm")
    xb = x.to(torch.bfloat16)                           # bf16 token stream input
    yb = ops.layernorm(xb, g, b, 1e-5)
    torch.cuda.synchronize()
    _close(yb, F.layer_norm(xb.float(), (C,), g, b, 1e-5), 1e-2, "layernorm bf16 in")


# ------------------------------------------------------------------------------------------------ attention
def _view_ref(qkv, heads, cross):
    Fr, H, V, w, C3 = qkv.shape
    C = C3 // 3
    q, k, v = qkv.float().split(C, dim=-1)
    out = torch.empty_like(q)

    d = C // heads

    def mha(qi, ki, vi):
        B, Nq, _ = qi.shape
        qh = qi.reshape(B, Nq, heads, d).transpose(1, 2)
        kh = ki.reshape(B, -1, heads, d).transpose(1, 2)
        vh = vi.reshape(B, -1, heads, d).transpose(1, 2)
        return F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(B, Nq, C)

    for i in range(V):
        nb = NEIGH[i] if cross else (i,)
        ki = torch.cat([k[:, :, j] for j in nb], dim=2).reshape(Fr, -1, C)
        vi = torch.cat([v[:, :, j] for j in nb], dim=2).reshape(Fr, -1, C)
        out[:, :, i] = mha(q[:, :, i].reshape(Fr, H * w, C), ki, vi).reshape(Fr, H, w, C)
    return out


@pytest.mark.parametrize("Fr,H,w,heads", [(2, 8, 16, 1), (1, 32, 56, 5), (2, 16, 28, 2), (2, 8, 14, 3), (3, 4, 7, 2),
                                          (1, 32, 64, 2), (2, 2, 3, 1), (1, 1, 2, 1), (2, 16, 24, 2), (2, 12, 28, 2), (1, 20, 28, 1)])
@pytest.mark.parametrize("cross", [False, True])
def test_attention_view(ops, Fr, H, w, heads, cross):
    C = heads * 64
    qkv = _rand((Fr, H, 6, w, 3 * C), 10, 1.0, torch.bfloat16)
    out = ops.attention_view(qkv, heads, cross, NEIGH)
    torch.cuda.synchronize()
    _close(out, _view_ref(qkv, heads, cross), 2e-2, f"attention_view cross={cross}")


@pytest.mark.parametrize("Fr,H,w,heads", [(1, 32, 56, 4), (2, 16, 28, 2), (2, 8, 14, 3), (1, 32, 64, 2), (2, 2, 3, 1), (2, 16, 24, 2), (2, 12, 28, 2), (1, 20, 28, 1)])
@pytest.mark.parametrize("cross", [False, True])
def test_attention_view_head_dim_80(ops, Fr, H, w, heads, cross):
    """BASELINE.json configs[4]: head_dim 80 = a 64-channel tile + a 16-channel tile per Q/K/V (fifth K step of S, second
    N = 16 MMA of P V), incl. the headline 32x56 views and the native 32x64 views (64-key blocks there)."""
    C = heads * 80
    qkv = _rand((Fr, H, 6, w, 3 * C), 18, 1.0, torch.bfloat16)
    out = ops.attention_view(qkv, heads, cross, NEIGH)
    torch.cuda.synchronize()
    _close(out, _view_ref(qkv, heads, cross), 2e-2, f"attention_view d=80 cross={cross}")


def test_attention_text_head_dim_80(ops):
    b, Nq, heads, Nk = 2, 700, 2, 77
    C = heads * 80
    q = _rand((b, Nq, C), 19, 1.0, torch.bfloat16)
    kv = _rand((b, Nk, 2 * C), 20, 1.0, torch.bfloat16)
    out = ops.attention_text(q, kv, heads)
    torch.cuda.synchronize()
    qh = q.float().reshape(b, Nq, heads, 80).transpose(1, 2)
    kh = kv.float()[..., :C].reshape(b, Nk, heads, 80).transpose(1, 2)
    vh = kv.float()[..., C:].reshape(b, Nk, heads, 80).transpose(1, 2)
    ref = F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(b, Nq, C)
    _close(out, ref, 2e-2, "attention_text d=80")


def test_attention_view_sharp_softmax(ops):
    """Large logits (peaked softmax) exercise the running-max rescale across K/V blocks."""
    Fr, H, w, heads = 1, 16, 28, 2
    C = heads * 64
    qkv = _rand((Fr, H, 6, w, 3 * C), 11, 3.0, torch.bfloat16)
    out = ops.attention_view(qkv, heads, True, NEIGH)
    torch.cuda.synchronize()
    _close(out, _view_ref(qkv, heads, True), 3e-2, "attention_view sharp")


@pytest.mark.parametrize("b,Nq,heads,Nk", [(2, 300, 2, 77), (1, 10752, 5, 77), (2, 128, 1, 77), (1, 50, 3, 16), (2, 1000, 2, 128)])
def test_attention_text(ops, b, Nq, heads, Nk):
    C = heads * 64
    q = _rand((b, Nq, C), 12, 1.0, torch.bfloat16)
    kv = _rand((b, Nk, 2 * C), 13, 1.0, torch.bfloat16)
    out = ops.attention_text(q, kv, heads)
    torch.cuda.synchronize()
    qh = q.float().reshape(b, Nq, heads, 64).transpose(1, 2)
    kh = kv.float()[..., :C].reshape(b, Nk, heads, 64).transpose(1, 2)
    vh = kv.float()[..., C:].reshape(b, Nk, heads, 64).transpose(1, 2)
    ref = F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(b, Nq, C)
    _close(out, ref, 2e-2, "attention_text")


@pytest.mark.parametrize("d", [64, 80])
@pytest.mark.parametrize("b,T,P,heads", [(2, 8, 100, 2), (1, 8, 2688, 5), (2, 4, 33, 1), (1, 16, 64, 2), (1, 1, 10, 1), (1, 5, 40, 2),
                                           (2, 12, 17, 1)])
def test_attention_temporal(ops, b, T, P, heads, d):
    C = heads * d
    qkv = _rand((b, T, P, 3 * C), 14, 1.0, torch.bfloat16)
    out = ops.attention_temporal(qkv, heads)
    torch.cuda.synchronize()
    q, k, v = qkv.float().split(C, dim=-1)

    def hs(z):  # [b,T,P,C] -> [b*P, heads, T, d]
        return z.permute(0, 2, 1, 3).reshape(b * P, T, heads, d).transpose(1, 2)

    ref = F.scaled_dot_product_attention(hs(q), hs(k), hs(v)).transpose(1, 2).reshape(b, P, T, C).permute(0, 2, 1, 3)
    _close(out, ref, 2e-2, "attention_temporal")


# ------------------------------------------------------------------------------------------------ convs / layout
@pytest.mark.parametrize("Fr,H,W,Cin,Cout,stride,silu", [(2, 16, 24, 8, 320, 1, False), (1, 32, 48, 20, 16, 1, True),
                                                        (2, 16, 24, 32, 96, 2, True), (1, 8, 12, 320, 4, 1, False),
                                                        (1, 9, 13, 96, 256, 2, True)])
def test_conv3x3_direct(ops, Fr, H, W, Cin, Cout, stride, silu):
    x = _rand((Fr, H, W, Cin), 15)
    w = _rand((Cout, Cin, 3, 3), 16, (9 * Cin) ** -0.5)
    bias = _rand((Cout,), 17, 0.1)
    cpad = (Cout + 15) // 16 * 16
    wp = torch.zeros(9, Cin, cpad, device="cuda")
    wp[:, :, :Cout] = w.permute(2, 3, 1, 0).reshape(9, Cin, Cout)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    addend = _rand((Fr, Ho, Wo, Cout), 18)
    y = ops.conv3x3_direct(x, wp, bias, Cout, stride=stride, silu=silu, addend=addend)
    torch.cuda.synchronize()
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, bias, stride=stride, padding=1)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 3, 1) + addend
    _close(y, ref, 1e-4, "conv3x3_direct")
    yb = ops.conv3x3_direct(x.to(torch.bfloat16), wp, bias, Cout, stride=stride, silu=silu, out_dtype=torch.bfloat16)
    torch.cuda.synchronize()
    ref2 = F.conv2d(x.to(torch.bfloat16).float().permute(0, 3, 1, 2), w, bias, stride=stride, padding=1)
    ref2 = (F.silu(ref2) if silu else ref2).permute(0, 2, 3, 1)
    _close(yb, ref2, 1e-2, "conv3x3_direct bf16")


def test_im2col_s2_matches_strided_conv(ops):
    Fr, H, W, C, N = 2, 16, 24, 64, 160
    x = _rand((Fr, H, W, C), 19)
    w = _rand((N, C, 3, 3), 20, (9 * C) ** -0.5)
    cols, (f, Ho, Wo) = ops.im2col_s2(x)
    wp = w.permute(0, 2, 3, 1).reshape(N, 9 * C).contiguous().to(torch.bfloat16)
    out = ops.gemm(cols, wp)
    torch.cuda.synchronize()
    ref = F.conv2d(x.to(torch.bfloat16).float().permute(0, 3, 1, 2), wp.float().reshape(N, 3, 3, C).permute(0, 3, 1, 2),
                   stride=2, padding=1).permute(0, 2, 3, 1)
    _close(out.reshape(Fr, Ho, Wo, N), ref, 2e-3, "im2col_s2 + gemm")
    # pad = 0: the VAE encoder's Downsample (zero row / column appended at the far edges, stride-2 conv without padding)
    cols0, (f, Ho0, Wo0) = ops.im2col_s2(x, pad=0)
    out0 = ops.gemm(cols0, wp)
    torch.cuda.synchronize()
    xp = F.pad(x.to(torch.bfloat16).float().permute(0, 3, 1, 2), (0, 1, 0, 1))
    ref0 = F.conv2d(xp, wp.float().reshape(N, 3, 3, C).permute(0, 3, 1, 2), stride=2, padding=0).permute(0, 2, 3, 1)
    assert (Ho0, Wo0) == (H // 2, W // 2)
    _close(out0.reshape(Fr, Ho0, Wo0, N), ref0, 2e-3, "im2col_s2(pad=0) + gemm")


def test_layout_and_small_ops(ops):
    x = _rand((3, 5, 12, 20), 21)
    nhwc = ops.nchw_to_nhwc(x)
    back = ops.nhwc_to_nchw(nhwc)
    wide = torch.zeros(3, 12, 20, 9, device="cuda")
    ops.nchw_to_nhwc(x, out=wide, ch_off=4)
    torch.cuda.synchronize()
    assert torch.equal(nhwc, x.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(back, x)
    assert torch.equal(wide[..., 4:], x.permute(0, 2, 3, 1)) and wide[..., :4].abs().max() == 0

    h = _rand((2, 6, 8, 64), 22)
    up = ops.upsample2x(h)
    ref = F.interpolate(h.permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    skip = _rand((2, 6, 8, 32), 23); ctrl = _rand((2, 6, 8, 32), 24)
    cat = ops.concat_add(h, skip, ctrl)
    cat2 = ops.concat_add(h, skip, None)
    a = h.clone(); ops.add_(a, h)
    cb = ops.cast_operand(h)
    torch.cuda.synchronize()
    assert torch.equal(up, ref.to(torch.bfloat16))
    assert torch.equal(cat, torch.cat([h, skip + ctrl], -1)) and torch.equal(cat2, torch.cat([h, skip], -1))
    assert torch.equal(a, 2 * h) and torch.equal(cb, h.to(torch.bfloat16))


def test_timestep_embedding_and_linear_small(ops):
    t = torch.tensor([0, 39, 500, 999, 999, 7], dtype=torch.int64, device="cuda")
    emb = ops.timestep_embedding(t, 320)
    half = 160
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device="cuda") / half)
    args = t[:, None].float() * freqs[None]
    ref = torch.cat([torch.cos(args), torch.sin(args)], -1)
    torch.cuda.synchronize()
    _close(emb, ref, 2e-4, "timestep_embedding")
    for M in (6, 16, 32):
        x = _rand((M, 320), 25)
        w = _rand((1280, 320), 26, 320 ** -0.5, torch.bfloat16)
        b = _rand((1280,), 27, 0.1)
        y = ops.linear_small(x, w, b, silu_in=True, silu_out=True)
        torch.cuda.synchronize()
        _close(y, F.silu(F.linear(F.silu(x), w.float(), b)), 1e-4, "linear_small")


def test_cfg_euler_and_scale_dup(ops):
    n = (8, 4, 32, 336)
    x = _rand(n, 28, 10.0)
    eps2 = _rand((16, 4, 32, 336), 29)
    x0 = x.clone()
    xin = torch.empty(16, 4, 32, 336, device="cuda")
    sigma, nxt, scale = 14.61464, 11.54277, 5.0
    c_in = 1.0 / math.sqrt(nxt * nxt + 1.0)
    ops.cfg_euler_step(x, eps2, xin, sigma, nxt, scale, c_in)
    torch.cuda.synchronize()
    den_u = eps2[:8] * (-sigma) + x0
    den_c = eps2[8:] * (-sigma) + x0
    den = den_u + scale * (den_c - den_u)
    ref = x0 + (nxt - sigma) * ((x0 - den) / sigma)
    _close(x, ref, 1e-5, "cfg_euler x")
    _close(xin, torch.cat([ref, ref]) * c_in, 1e-5, "cfg_euler x_in")
    d = ops.scale_dup(x0, 0.25, 2)
    torch.cuda.synchronize()
    assert torch.equal(d, torch.cat([x0 * 0.25, x0 * 0.25]))
