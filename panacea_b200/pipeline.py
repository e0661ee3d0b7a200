"""Sampling glue around the hot path: the device-side part of `DiffusionEngine3D.sample`
(reference sgm/models/diffusion.py:233-255) and a builder that instantiates network + denoiser + sampler from
the reference YAML's `model.params` block (configs/inference_nuscenes.yaml)."""
from __future__ import annotations

import torch

from .sgm.modules.diffusionmodules.sampling import BoundDenoiser
from .sgm.modules.diffusionmodules.wrappers import OpenAIWrapperControlLDM3D
from .sgm.util import instantiate_from_config

DEFAULT_DENOISER = {
    "target": "sgm.modules.diffusionmodules.denoiser.DiscreteDenoiser",
    "params": {
        "num_idx": 1000,
        "weighting_config": {"target": "sgm.modules.diffusionmodules.denoiser_weighting.EpsWeighting"},
        "scaling_config": {"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"},
        "discretization_config": {"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"},
    },
}


def default_sampler_config(num_steps: int = 25, scale: float = 5.0) -> dict:
    return {
        "target": "sgm.modules.diffusionmodules.sampling.EulerEDMSampler",
        "params": {
            "num_steps": num_steps,
            "discretization_config": {"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"},
            "guider_config": {"target": "sgm.modules.diffusionmodules.guiders.VanillaCFG", "params": {"scale": scale}},
        },
    }


def default_network_config(**over) -> dict:
    """network_config of configs/inference_nuscenes.yaml:30-71 with anchors resolved."""
    base = dict(insert_crossview=True, spatial_only_attn_type="intra-view", use_checkpoint=True, use_fp16=True,
                in_channels=8, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2,
                channel_mult=[1, 2, 4, 4], num_head_channels=64, use_spatial_transformer=True,
                use_linear_in_transformer=True, transformer_depth=1, context_dim=1024, legacy=False, num_frames=8, alpha=1)
    base.update(over)
    cn = {k: v for k, v in base.items() if k not in ("out_channels", "use_fp16")}
    cn.update(hint_channels=19, control_scales=1.0)
    return {"target": "sgm.modules.diffusionmodules.controlmodel.ControlledUNetModel3D",
            "params": dict(base, controlnet_config={"target": "sgm.modules.diffusionmodules.controlmodel.ControlNet3D", "params": cn})}


class DenoisingPipeline:
    """network wrapper + denoiser + sampler, i.e. the members `DiffusionEngine3D.__init__` builds for the hot path
    (diffusion.py:65-95) minus conditioner / first stage (out of scope)."""

    def __init__(self, network_config=None, denoiser_config=None, sampler_config=None, use_cuda_graph=True,
                 share_noise_level: float = 0.07, precision: str | None = None):
        self.model = instantiate_from_config(network_config or default_network_config())
        if precision is not None:
            self.model.set_precision(precision)
        self.wrapper = OpenAIWrapperControlLDM3D(self.model, use_cuda_graph=use_cuda_graph)
        self.denoiser = instantiate_from_config(denoiser_config or DEFAULT_DENOISER)
        self.sampler = instantiate_from_config(sampler_config or default_sampler_config())
        self.share_noise_level = share_noise_level
        self.num_frames = self.model.cfg.num_frames

    def to(self, device):
        self.wrapper.to(device)
        return self

    @torch.no_grad()
    def sample(self, cond: dict, uc: dict, randn: torch.Tensor, num_steps=None, share_noise: bool = True):
        """diffusion.py:242-254: the shared-noise mix `randn += share_noise_level * repeat(concat[-1], t=num_frames)`,
        applied like the reference whenever share_noise_level > 0 (it is 0.07 in inference_nuscenes.yaml; with
        use_last_frame data concat[-1] is the conditioning frame's latent), then the sampler. `randn` is the caller's
        CPU-generator draw moved to the device (the reference draws on CPU, :242). share_noise=False opts out."""
        x = randn.float()
        if share_noise and self.share_noise_level > 0.0:
            x = x + self.share_noise_level * cond["concat"][-1:].float().expand_as(x)
        return self.sampler(BoundDenoiser(self.denoiser, self.wrapper), x, cond, uc, num_steps=num_steps)
