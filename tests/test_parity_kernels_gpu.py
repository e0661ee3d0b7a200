"""GPU tests of the parity-mode building blocks (panacea_b200.ops.ParityOps) against float64 torch math:
split-bf16 operands through the tcgen05 GEMM / implicit conv, the fp32 attention kernels (head_dim 64 and 80), the
exact-erf GEGLU pass and the split stores of the normalisation kernels. Tolerances are fp32-class (1e-5 .. 1e-4)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

NEIGH = ((5, 1), (0, 2), (1, 3), (2, 4), (3, 5), (4,))


@pytest.fixture(scope="module")
def ops():
    from panacea_b200.ops import ParityOps
    return ParityOps()


def _rand(shape, seed, scale=1.0, shift=0.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale + shift).cuda()


def _dec(y):
    """[hi | lo | hi] -> fp32 value."""
    C = y.shape[-1] // 3
    assert torch.equal(y[..., :C], y[..., 2 * C:])
    return y[..., :C].float() + y[..., C:2 * C].float()


def _rel(got, ref):
    return ((got.double() - ref.double()).norm() / ref.double().norm()).item()


@pytest.mark.parametrize("M,K,N", [(300, 320, 960), (4096, 1280, 320), (172032 // 8, 320, 2560), (77, 1024, 640)])
def test_split3_gemm_is_fp32_class(ops, M, K, N):
    from panacea_b200.ops import split3
    a = _rand((M, K), 1)
    w = _rand((N, K), 2, K ** -0.5)
    bias = _rand((N,), 3)
    res = _rand((M, N), 4)
    a_op = ops.cast_operand(a)
    assert a_op.shape == (M, 3 * K) and _rel(_dec(a_op), a) < 1e-5
    y = ops.gemm(a_op, split3(w), bias=bias, residual=res)
    torch.cuda.synchronize()
    ref = a.double() @ w.double().t() + bias.double() + res.double()
    assert _rel(y, ref) < 2e-5, _rel(y, ref)


@pytest.mark.parametrize("NB,H,W,C,N,taps", [(2, 16, 48, 64, 160, (3, 3)), (4, 32, 56, 320, 320, (3, 3)), (2, 8, 96, 128, 128, (3, 1))])
def test_split3_implicit_conv_is_fp32_class(ops, NB, H, W, C, N, taps):
    from panacea_b200.ops import split3
    th, tw = taps
    x = _rand((NB, H, W, C), 5)
    w = _rand((N, th * tw * C), 6, (th * tw * C) ** -0.5)
    y = ops.gemm(ops.cast_operand(x), split3(w, th * tw), taps=taps)
    torch.cuda.synchronize()
    wk = w.double().reshape(N, th, tw, C).permute(0, 3, 1, 2)
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), wk, padding=(th // 2, tw // 2)).permute(0, 2, 3, 1)
    assert _rel(y.reshape(ref.shape), ref) < 2e-5


def _mha64(q, k, v, heads):
    B, Nq, C = q.shape
    d = C // heads
    qh = q.double().reshape(B, Nq, heads, d).transpose(1, 2)
    kh = k.double().reshape(B, -1, heads, d).transpose(1, 2)
    vh = v.double().reshape(B, -1, heads, d).transpose(1, 2)
    s = (qh @ kh.transpose(-1, -2)) * (d ** -0.5)
    return (s.softmax(-1) @ vh).transpose(1, 2).reshape(B, Nq, C)


@pytest.mark.parametrize("d", [64, 80])
@pytest.mark.parametrize("cross", [False, True])
def test_attention_view_f32(ops, d, cross):
    Fr, H, V, w, heads = 2, 8, 6, 14, 2
    C = heads * d
    qkv = _rand((Fr, H, V, w, 3 * C), 7)
    out = _dec(ops.attention_view(qkv, heads, cross, NEIGH))
    torch.cuda.synchronize()
    q, k, v = qkv.split(C, dim=-1)
    for i in range(V):
        nb = NEIGH[i] if cross else (i,)
        ki = torch.cat([k[:, :, j] for j in nb], dim=2).reshape(Fr, -1, C)
        vi = torch.cat([v[:, :, j] for j in nb], dim=2).reshape(Fr, -1, C)
        ref = _mha64(q[:, :, i].reshape(Fr, H * w, C), ki, vi, heads).reshape(Fr, H, w, C)
        assert _rel(out[:, :, i], ref) < 1e-5


@pytest.mark.parametrize("d", [64, 80])
def test_attention_text_and_temporal_f32(ops, d):
    heads = 3
    C = heads * d
    q = _rand((2, 500, C), 8)
    kv = _rand((2, 77, 2 * C), 9)
    out = _dec(ops.attention_text(q, kv, heads))
    ref = _mha64(q, kv[..., :C], kv[..., C:], heads)
    assert _rel(out, ref) < 1e-5
    for T in (1, 4, 8, 16):
        b, P = 2, 37
        qkv = _rand((b, T, P, 3 * C), 10 + T)
        o = _dec(ops.attention_temporal(qkv, heads))
        qq, kk, vv = qkv.split(C, dim=-1)
        seq = lambda z: z.permute(0, 2, 1, 3).reshape(b * P, T, C)
        r = _mha64(seq(qq), seq(kk), seq(vv), heads).reshape(b, P, T, C).permute(0, 2, 1, 3)
        assert _rel(o, r) < 1e-5


def test_geglu_pass_uses_the_exact_erf(ops):
    from panacea_b200.ops import geglu_pack, split3
    M, K, inner = 1000, 320, 1280
    a = _rand((M, K), 20)
    w = _rand((2 * inner, K), 21, K ** -0.5)
    b = _rand((2 * inner,), 22)
    y = _dec(ops.gemm(ops.cast_operand(a), split3(geglu_pack(w)), bias=geglu_pack(b).contiguous(), geglu=True))
    torch.cuda.synchronize()
    h = a.double() @ w.double().t() + b.double()
    ref = h[:, :inner] * F.gelu(h[:, inner:])
    assert _rel(y, ref) < 2e-5


def test_norm_kernels_store_split_operands(ops):
    x = _rand((3, 700, 320), 30, 2.0, 0.5)
    g = _rand((320,), 31, 0.1, 1.0); b = _rand((320,), 32, 0.1)
    y, raw = ops.groupnorm(x, g, b, 1e-5, True, want_raw=True)
    ref = F.silu(F.group_norm(x.double().permute(0, 2, 1), 32, g.double(), b.double(), 1e-5)).permute(0, 2, 1)
    assert _rel(_dec(y), ref) < 1e-5 and _rel(_dec(raw), x) < 1e-5
    yf = ops.groupnorm(x, g, b, 1e-5, True, out_f32=True)
    assert yf.dtype == torch.float32 and _rel(yf, ref) < 1e-5
    xp = _rand((2, 8, 50, 640), 33, 1.5, -0.3)
    gp = _rand((640,), 34, 0.1, 1.0); bp = _rand((640,), 35, 0.1)
    yp = _dec(ops.groupnorm_pixel(xp, gp, bp, 1e-5, True))
    z = xp.double().permute(0, 2, 3, 1).reshape(100, 640, 8)
    rp = F.silu(F.group_norm(z, 32, gp.double(), bp.double(), 1e-5)).reshape(2, 50, 640, 8).permute(0, 3, 1, 2)
    assert _rel(yp, rp) < 1e-5
    xl = _rand((999, 1280), 36, 3.0, 1.0)
    gl = _rand((1280,), 37, 0.1, 1.0); bl = _rand((1280,), 38, 0.1)
    yl = _dec(ops.layernorm(xl, gl, bl))
    assert _rel(yl, F.layer_norm(xl.double(), (1280,), gl.double(), bl.double(), 1e-5)) < 1e-5
    xu = _rand((2, 4, 6, 64), 39)
    assert _rel(_dec(ops.upsample2x(xu)), xu.repeat_interleave(2, 1).repeat_interleave(2, 2)) < 1e-5
    cols, (Fr, Ho, Wo) = ops.im2col_s2(xu)
    assert cols.shape == (2 * 2 * 3, 9 * 3 * 64)


def test_groupnorm_two_phase_form_matches_the_fused_one():
    """PN_GN_TWO_PHASE=1 (devices that cannot hold the cooperative grid) must give the same bits as the fused launch."""
    import os, subprocess, sys
    code = ("import torch;from panacea_b200.ops import NativeOps;o=NativeOps();g=torch.Generator().manual_seed(0);"
            "x=torch.randn(4,3000,640,generator=g).cuda();w=torch.ones(640).cuda();b=torch.zeros(640).cuda();"
            "y=o.groupnorm(x,w,b,1e-5,True);torch.cuda.synchronize();print(float(y.float().double().sum()),float(y.float().abs().double().sum()))")
    outs = []
    for flag in ("0", "1"):
        env = dict(os.environ, PN_GN_TWO_PHASE=flag)
        outs.append(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True,
                                   cwd=str(__import__("pathlib").Path(__file__).resolve().parent.parent)).stdout.strip())
    assert outs[0] == outs[1], outs
