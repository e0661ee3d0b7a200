"""Conditioner glue around the hot path: `GeneralConditioner`
(reference sgm/modules/encoders/modules.py:95-220) with the three embedders configs/inference_nuscenes.yaml:73-93 names.

The conditioner runs ONCE per sample, before the denoising loop; it only routes tensors (rearrange, cat, zeros for the
unconditional branch) and is therefore plain host/torch code, like the reference's. The learned embedders are OUT OF
SCOPE of this repository (BASELINE.json configs[3]: "random-init VAE/CLIP stubs"): `FrozenOpenCLIPEmbedder` here is a
deterministic stand-in that maps each prompt string to a reproducible [77, 1024] tensor — there are no OpenCLIP weights in
this environment — and `VAEEmbedder` delegates to whatever first-stage model the engine hands it (a stub or a real one).
"""
from __future__ import annotations

import hashlib
from contextlib import nullcontext

import torch
import torch.nn as nn

from ...util import instantiate_from_config


class AbstractEmbModel(nn.Module):
    """modules.py:51-92: carries is_trainable / ucg_rate / input_key."""

    def __init__(self):
        super().__init__()
        self.is_trainable = False
        self.ucg_rate = 0.0
        self.input_key = None
        self.legacy_ucg_val = None


class IdentityEncoder(AbstractEmbModel):
    """modules.py:244-249: the BEV control maps pass through unchanged (they become c["cond_feat"])."""

    def encode(self, x):
        return x

    def forward(self, x):
        return x


class FrozenOpenCLIPEmbedder(AbstractEmbModel):
    """Stand-in for modules.py:555-640 (OpenCLIP ViT-H/14 text tower, penultimate layer, [b, 77, 1024]). The real tower
    and its weights are not available here and are outside the hot path; this stub keeps the interface and is
    deterministic per prompt string, so the conditional ("a driving scene ...") and unconditional ("") branches differ
    reproducibly on every rank."""

    def __init__(self, arch="ViT-H-14", version="laion2b_s32b_b79k", device="cuda", max_length=77, freeze=True,
                 layer="penultimate", always_return_pooled=False, legacy=True, context_dim=1024):
        super().__init__()
        self.max_length, self.context_dim = max_length, context_dim
        self.register_buffer("_dev", torch.zeros(1), persistent=False)

    def encode(self, text):
        return self(text)

    @torch.no_grad()
    def forward(self, text):
        outs = []
        for s in text:
            seed = int.from_bytes(hashlib.sha256(str(s).encode()).digest()[:8], "little") % (2 ** 63)
            g = torch.Generator().manual_seed(seed)
            outs.append(torch.randn(self.max_length, self.context_dim, generator=g))
        return torch.stack(outs).to(self._dev.device)


class VAEEmbedder(AbstractEmbModel):
    """modules.py:1000-1055: encodes the image-condition frames with the engine's first stage and scales the latent
    (the engine injects first_stage_model / scale_factor / disable_first_stage_autocast, diffusion.py:111-122)."""

    def __init__(self, down_blur_factor: int = 1):
        super().__init__()
        if down_blur_factor != 1:
            raise NotImplementedError("down_blur_factor > 1 is not used by the reference config")
        self.first_stage_model = None
        self.scale_factor = None

    def freeze(self):
        return self

    def encode(self, x):
        return self(x)

    @torch.no_grad()
    def forward(self, x):
        assert self.first_stage_model is not None and self.scale_factor is not None, "first_stage_model / scale_factor not set"
        return self.scale_factor * self.first_stage_model.encode(x)


class GeneralConditioner(nn.Module):
    """modules.py:95-220, same routing rules: output key by tensor rank (2 vector, 3 crossattn, 4/5 concat), the
    `cond_img` embedder feeds "cond_feat", [b t c h w] inputs are flattened to (b t), equal keys are concatenated,
    `force_zero_embeddings` zeroes an embedder's output for the unconditional branch."""
    OUTPUT_DIM2KEYS = {2: "vector", 3: "crossattn", 4: "concat", 5: "concat"}
    KEY2CATDIM = {"vector": 1, "crossattn": 2, "concat": 1}

    def __init__(self, emb_models):
        super().__init__()
        embedders = []
        for embconfig in emb_models:
            embedder = instantiate_from_config(embconfig)
            assert isinstance(embedder, AbstractEmbModel), f"{type(embedder).__name__} has to inherit from AbstractEmbModel"
            embedder.is_trainable = embconfig.get("is_trainable", False)
            embedder.ucg_rate = embconfig.get("ucg_rate", 0.0)
            if "input_key" in embconfig:
                embedder.input_key = embconfig["input_key"]
            elif "input_keys" in embconfig:
                embedder.input_keys = embconfig["input_keys"]
            else:
                raise KeyError(f"need either 'input_key' or 'input_keys' for embedder {type(embedder).__name__}")
            embedder.legacy_ucg_val = embconfig.get("legacy_ucg_value", None)
            if embedder.legacy_ucg_val is not None:
                raise NotImplementedError("legacy_ucg_value (training-time dropout) is not part of the inference path")
            embedders.append(embedder.eval())
        self.embedders = nn.ModuleList(embedders)

    def forward(self, batch: dict, force_zero_embeddings=None) -> dict:
        output = {}
        force_zero_embeddings = force_zero_embeddings or []
        for embedder in self.embedders:
            with (nullcontext() if embedder.is_trainable else torch.no_grad()):
                if getattr(embedder, "input_key", None) is not None:
                    x = batch[embedder.input_key]
                    if embedder.input_key in ("final_cond_zero", "cond_img"):      # modules.py:157-166
                        x = x.reshape(-1, *x.shape[2:]).contiguous()               # "b t c h w -> (b t) c h w"
                    emb_out = embedder(x)
                else:
                    emb_out = embedder(*[batch[k] for k in embedder.input_keys])
            if not isinstance(emb_out, (list, tuple)):
                emb_out = [emb_out]
            for emb in emb_out:
                out_key = "cond_feat" if getattr(embedder, "input_key", None) == "cond_img" else self.OUTPUT_DIM2KEYS[emb.dim()]
                if embedder.ucg_rate > 0.0:                                         # modules.py:184-195 (training-time dropout)
                    keep = torch.bernoulli((1.0 - embedder.ucg_rate) * torch.ones(emb.shape[0], device=emb.device))
                    emb = keep.reshape(-1, *([1] * (emb.dim() - 1))) * emb
                if getattr(embedder, "input_key", None) in force_zero_embeddings:
                    emb = torch.zeros_like(emb)
                if out_key in output:
                    output[out_key] = torch.cat((output[out_key], emb), self.KEY2CATDIM[out_key])
                else:
                    output[out_key] = emb
        return output

    def get_unconditional_conditioning(self, batch_c, batch_uc=None, force_uc_zero_embeddings=None):
        """modules.py:204-219: ucg dropout disabled for both passes."""
        rates = [e.ucg_rate for e in self.embedders]
        for e in self.embedders:
            e.ucg_rate = 0.0
        c = self(batch_c)
        uc = self(batch_c if batch_uc is None else batch_uc, force_uc_zero_embeddings or [])
        for e, r in zip(self.embedders, rates):
            e.ucg_rate = r
        return c, uc
