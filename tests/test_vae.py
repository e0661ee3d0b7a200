"""N2: the VAE decoder on the hot path's kernels, against the output of the UNMODIFIED reference
`Decoder(post_quant_conv(z))` (sgm/modules/diffusionmodules/model.py:882-1030, autoencoder.py:362-365) on a shrunk
ddconfig with seeded weights (tests/golden/vae_decode_small.pt, made by oracle/make_golden.py --only vae)."""
from pathlib import Path

import pytest
import torch

from oracle.make_golden import VAE_DDCONFIG, vae_decoder_input, vae_decoder_weights, vae_encoder_input
from panacea_b200.vae import VAEDecoderEngine, VAEEncoderEngine, decoder_param_spec, encoder_param_spec

GOLDEN = Path(__file__).resolve().parent / "golden"


def test_decoder_orchestration_matches_the_reference_on_cpu():
    from torch_ref_ops import TorchRefOps
    g = torch.load(GOLDEN / "vae_decode_small.pt")
    eng = VAEDecoderEngine(VAE_DDCONFIG, TorchRefOps())
    assert sorted(eng.spec) == g["keys"]
    eng.pack(vae_decoder_weights(eng.spec))
    out = eng.decode(vae_decoder_input())
    assert out.shape == g["image"].shape and (out - g["image"]).abs().max().item() < 5e-5


def test_encoder_orchestration_matches_the_reference_on_cpu():
    """quant_conv(Encoder(x)) (model.py:763-880; autoencoder.py:352-357), incl. the asymmetric (0,1,0,1) padding of the
    stride-2 Downsample convs."""
    from torch_ref_ops import TorchRefOps
    g = torch.load(GOLDEN / "vae_decode_small.pt")
    eng = VAEEncoderEngine(VAE_DDCONFIG, TorchRefOps())
    assert sorted(eng.spec) == g["encoder_keys"]
    eng.pack(vae_decoder_weights(eng.spec, seed=9))
    out = eng.encode_moments(vae_encoder_input())
    assert out.shape == g["moments"].shape and (out - g["moments"]).abs().max().item() < 5e-5


def test_full_size_spec_is_the_sd_vae_decoder():
    dd = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4], num_res_blocks=2,
              attn_resolutions=[], dropout=0.0)            # configs/inference_nuscenes.yaml:98-110
    spec = decoder_param_spec(dd, 4)
    n = sum(int(torch.tensor(s).prod()) for s in spec.values())
    assert spec["decoder.conv_in.weight"] == (512, 4, 3, 3) and spec["decoder.up.1.block.0.nin_shortcut.weight"] == (256, 512, 1, 1)
    assert spec["decoder.conv_out.weight"] == (3, 128, 3, 3) and "decoder.up.0.upsample.conv.weight" not in spec
    assert 49_400_000 < n < 49_600_000                     # 49.5 M parameters: the SD VAE decoder + post_quant_conv


def test_mirror_first_stage_loads_reference_keys():
    from panacea_b200.sgm.models.autoencoder import AutoencoderKLInferenceWrapper
    m = AutoencoderKLInferenceWrapper(embed_dim=4, ddconfig=VAE_DDCONFIG, lossconfig={"target": "torch.nn.Identity"})
    sd = vae_decoder_weights(decoder_param_spec(VAE_DDCONFIG, 4))
    sd["encoder.conv_in.weight"] = torch.zeros(1)            # encoder keys of a real checkpoint are ignored (strict=False)
    res = m.load_state_dict({"first_stage_model." + k: v for k, v in sd.items()}, strict=False)
    assert len(res.missing_keys) > 0                          # prefix mismatch on purpose: nothing matched
    res = m.load_state_dict(sd, strict=False)
    assert not [k for k in res.missing_keys if k.startswith(("decoder.", "post_quant_conv."))]
    assert torch.equal(m.state_dict()["decoder.conv_out.weight"], sd["decoder.conv_out.weight"])
    with pytest.raises(RuntimeError):
        m.decode(torch.zeros(2, 4, 8, 48))                    # no CPU path
    with pytest.raises(RuntimeError):
        m.encode(torch.zeros(2, 3, 32, 192))


@pytest.mark.gpu
def test_decoder_on_gpu_matches_the_reference():
    from panacea_b200.sgm.models.autoencoder import AutoencoderKLInferenceWrapper
    g = torch.load(GOLDEN / "vae_decode_small.pt")
    m = AutoencoderKLInferenceWrapper(embed_dim=4, ddconfig=VAE_DDCONFIG, lossconfig={"target": "torch.nn.Identity"})
    m.load_state_dict(vae_decoder_weights(decoder_param_spec(VAE_DDCONFIG, 4)), strict=False)
    m = m.cuda()
    out = m.decode(vae_decoder_input().cuda()).cpu()
    rel = ((out - g["image"]).norm() / g["image"].norm()).item()
    print(f"PARITY vae_decode_small rel_l2 {rel:.3e}")
    assert out.shape == g["image"].shape and rel < 1.5e-2     # bf16 operands, fp32 accumulation / norms / residuals


@pytest.mark.gpu
def test_encoder_on_gpu_matches_the_reference():
    from panacea_b200.sgm.models.autoencoder import AutoencoderKLInferenceWrapper
    g = torch.load(GOLDEN / "vae_decode_small.pt")
    m = AutoencoderKLInferenceWrapper(embed_dim=4, ddconfig=VAE_DDCONFIG, lossconfig={"target": "torch.nn.Identity"})
    m.load_state_dict(vae_decoder_weights(encoder_param_spec(VAE_DDCONFIG, 4), seed=9), strict=False)
    m = m.cuda()
    mom = m.encode_moments(vae_encoder_input().cuda()).cpu()
    rel = ((mom - g["moments"]).norm() / g["moments"].norm()).item()
    print(f"PARITY vae_encode_small rel_l2 {rel:.3e}")
    assert mom.shape == g["moments"].shape and rel < 1.5e-2
    torch.manual_seed(0)
    z = m.encode(vae_encoder_input().cuda())
    assert z.shape == (2, 4, 8, 48) and torch.isfinite(z).all()
    m.sample_posterior = False
    assert torch.equal(m.encode(vae_encoder_input().cuda()).cpu(), mom[:, :4])


@pytest.mark.gpu
def test_softmax_rows_kernel():
    from panacea_b200.ops import NativeOps
    ops = NativeOps()
    g = torch.Generator().manual_seed(3)
    for rows, N in ((7, 768), (300, 10752), (5, 64)):
        s = (torch.randn(rows, N, generator=g) * 20).cuda()
        p = ops.softmax_rows(s, 512 ** -0.5).float()
        ref = torch.softmax(s * 512 ** -0.5, dim=-1)
        assert (p - ref).abs().max().item() <= 4e-3 * ref.max().item() + 1e-6
