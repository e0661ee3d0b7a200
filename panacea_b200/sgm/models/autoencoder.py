"""First stage (VAE) stand-in behind the reference's interface (sgm/models/autoencoder.py:333-373,
`AutoencoderKLInferenceWrapper.encode / decode`, `post_quant_conv`). The KL autoencoder is row N2 of SURVEY.md
section 8f (out of the hot path; BASELINE.json configs[3] allows a random-init stub): this class keeps the tensor contract
— images [n, 3, 8h, 8w] in [-1, 1]  <->  latents [n, 4, h, w] — with a random-init 8x8 patch projection, so the engine
glue, the gather of decoded frames and the frame writers can run end to end. It is NOT a trained VAE."""
from __future__ import annotations

import torch
import torch.nn as nn


class AutoencoderKLInferenceWrapper(nn.Module):
    def __init__(self, embed_dim=4, ddconfig=None, lossconfig=None, monitor=None, seed: int = 1234, **kwargs):
        super().__init__()
        dd = ddconfig or {}
        self.in_channels = dd.get("in_channels", 3)
        self.z_channels = dd.get("z_channels", embed_dim)
        self.factor = 2 ** (len(dd.get("ch_mult", [1, 2, 4, 4])) - 1)           # 8x spatial compression
        g = torch.Generator().manual_seed(seed)
        f, ci, cz = self.factor, self.in_channels, self.z_channels
        self.quant_conv = nn.Conv2d(ci, cz, f, stride=f)                         # patch projection (mean of the posterior)
        self.post_quant_conv = nn.Conv2d(cz, cz, 1)
        self.decoder = nn.ConvTranspose2d(cz, ci, f, stride=f)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * (1.0 / max(p[0].numel(), 1)) ** 0.5)

    @torch.no_grad()
    def encode(self, x):
        """autoencoder.py:333-350 returns a posterior sample; the stub returns its mean (deterministic)."""
        return self.quant_conv(x.float())

    @torch.no_grad()
    def decode(self, z):
        return torch.tanh(self.decoder(self.post_quant_conv(z.float())))
