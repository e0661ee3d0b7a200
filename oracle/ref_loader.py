"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference (wenyuqing/panacea) hot-path modules from
/root/reference so the CPU restatement in oracle/unet_port.py can be pinned against them and golden
vectors can be generated (oracle/make_golden.py). Nothing here is product code, and nothing here works
on the GPU box (the reference tree does not travel) — only this container can call it.

The reference does not import out of the box: sgm/__init__.py pulls
pytorch_lightning, kornia, open_clip, omegaconf and xformers, none of which are installed. We inject
minimal stand-ins into sys.modules *before* importing sgm. The only arithmetic stand-in is
xformers.ops.memory_efficient_attention (third-party, xformers==0.0.16 pinned in requirements/pt13.txt:40;
call sites sgm/modules/attention.py:469-471,590-592): its published semantics are
softmax(q k^T / sqrt(d)) v on [B*heads, N, d] with no mask/bias, restated here with torch SDPA.
"""
from __future__ import annotations

import contextlib
import io
import math
import sys
import types
from pathlib import Path

import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = Path("/root/reference")


def reference_available() -> bool:
    return (REFERENCE_ROOT / "sgm" / "modules" / "attention.py").exists()


def _install_stubs() -> None:
    if "xformers" in sys.modules and getattr(sys.modules["xformers"], "_pn_stub", False):
        return

    pl = types.ModuleType("pytorch_lightning")
    pl.LightningModule = nn.Module
    pl.LightningDataModule = object
    sys.modules["pytorch_lightning"] = pl

    oc = types.ModuleType("omegaconf")

    class ListConfig(list):
        pass

    class DictConfig(dict):
        pass

    class OmegaConf(dict):
        @staticmethod
        def create(x):
            return x

    oc.ListConfig, oc.DictConfig, oc.OmegaConf = ListConfig, DictConfig, OmegaConf
    lc = types.ModuleType("omegaconf.listconfig")
    lc.ListConfig = ListConfig
    oc.listconfig = lc
    sys.modules["omegaconf"] = oc
    sys.modules["omegaconf.listconfig"] = lc

    for name in ("open_clip", "kornia"):
        sys.modules[name] = types.ModuleType(name)

    xf = types.ModuleType("xformers")
    xf._pn_stub = True
    xops = types.ModuleType("xformers.ops")

    def memory_efficient_attention(q, k, v, attn_bias=None, op=None):
        assert attn_bias is None
        return F.scaled_dot_product_attention(q, k, v)

    xops.memory_efficient_attention = memory_efficient_attention
    xf.ops = xops
    sys.modules["xformers"] = xf
    sys.modules["xformers.ops"] = xops


_sgm = None


def import_reference():
    """Returns a namespace with the reference's hot-path modules (openaimodel, controlmodel, attention, ...)."""
    global _sgm
    if _sgm is not None:
        return _sgm
    if not reference_available():
        raise RuntimeError("/root/reference is not present (it never is on the GPU box)")
    _install_stubs()
    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))
    with contextlib.redirect_stdout(io.StringIO()):
        import sgm  # noqa: F401
        from sgm.modules import attention
        from sgm.modules.diffusionmodules import (controlmodel, denoiser, denoiser_scaling, discretizer, guiders,
                                                  openaimodel, sampling, sampling_utils, util, wrappers)
    ns = types.SimpleNamespace(attention=attention, controlmodel=controlmodel, openaimodel=openaimodel,
                               wrappers=wrappers, util=util, denoiser=denoiser, denoiser_scaling=denoiser_scaling,
                               discretizer=discretizer, guiders=guiders, sampling=sampling,
                               sampling_utils=sampling_utils)
    _sgm = ns
    return ns


class _HShim:
    """Stands in for the `math` name inside sgm.modules.attention so that
    `H=int(math.sqrt(x.shape[1]//12))` (attention.py:428,537) returns the true latent height for views that
    are not 2:1 (e.g. 32x56). Everything else forwards to the real math module."""

    def __init__(self):
        self.table: dict[int, int] = {}

    def register(self, H: int, w: int) -> None:
        self.table[(H * 6 * w) // 12] = H

    def sqrt(self, v):
        if v in self.table:
            return float(self.table[v])
        return math.sqrt(v)

    def __getattr__(self, name):
        return getattr(math, name)


@contextlib.contextmanager
def view_height_shim(H: int, w: int, levels: int = 4):
    """Context manager enabling non-2:1 view shapes in the reference."""
    ref = import_reference()
    shim = _HShim()
    h, ww = H, w
    for _ in range(levels):
        shim.register(h, ww)
        h, ww = max(h // 2, 1), max(ww // 2, 1)
    old = ref.attention.math
    ref.attention.math = shim
    try:
        yield
    finally:
        ref.attention.math = old


def default_unet_kwargs(**over) -> dict:
    """configs/inference_nuscenes.yaml:30-71 as plain Python (anchors resolved)."""
    base = dict(
        insert_crossview=True, spatial_only_attn_type="intra-view", use_checkpoint=True, use_fp16=True,
        in_channels=8, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2,
        channel_mult=[1, 2, 4, 4], num_head_channels=64, use_spatial_transformer=True,
        use_linear_in_transformer=True, transformer_depth=1, context_dim=1024, legacy=False, num_frames=8, alpha=1,
    )
    base.update(over)
    return base


def default_controlnet_params(unet_kwargs: dict, hint_channels: int = 19) -> dict:
    keys = ["insert_crossview", "spatial_only_attn_type", "use_checkpoint", "in_channels", "model_channels",
            "attention_resolutions", "num_res_blocks", "channel_mult", "num_head_channels",
            "use_spatial_transformer", "use_linear_in_transformer", "transformer_depth", "context_dim", "legacy",
            "alpha", "num_frames"]
    p = {k: unet_kwargs[k] for k in keys}
    p.update(hint_channels=hint_channels, control_scales=1.0)
    return p


def build_reference_model(unet_kwargs: dict | None = None, hint_channels: int = 19):
    """Constructs the reference ControlledUNetModel3D (+ .controlnet) wrapped in OpenAIWrapperControlLDM3D."""
    ref = import_reference()
    kw = default_unet_kwargs() if unet_kwargs is None else dict(unet_kwargs)
    cn_cfg = {"target": "sgm.modules.diffusionmodules.controlmodel.ControlNet3D",
              "params": default_controlnet_params(kw, hint_channels)}
    with contextlib.redirect_stdout(io.StringIO()):
        model = ref.controlmodel.ControlledUNetModel3D(controlnet_config=cn_cfg, **kw)
        wrapper = ref.wrappers.OpenAIWrapperControlLDM3D(model, compile_model=False)
    return wrapper.eval()
