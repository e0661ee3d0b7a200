"""Mirror of the reference's `sgm` package for the denoising hot path only.

`panacea_b200.sgm.modules.diffusionmodules.*` exposes the same class names, constructor keywords, call
signatures and state-dict keys as the reference modules of the same dotted path (minus the leading
`panacea_b200.`), so `configs/inference_nuscenes.yaml` works by retargeting
`sgm.modules.diffusionmodules...` -> `panacea_b200.sgm.modules.diffusionmodules...`, or unchanged after
`panacea_b200.sgm.install_as_sgm()` when the reference package itself is not importable.
"""
import importlib
import sys

_MIRRORED = (
    "sgm", "sgm.util", "sgm.modules", "sgm.modules.diffusionmodules",
    "sgm.modules.diffusionmodules.controlmodel", "sgm.modules.diffusionmodules.wrappers",
    "sgm.modules.diffusionmodules.denoiser", "sgm.modules.diffusionmodules.denoiser_scaling",
    "sgm.modules.diffusionmodules.denoiser_weighting", "sgm.modules.diffusionmodules.discretizer",
    "sgm.modules.diffusionmodules.guiders", "sgm.modules.diffusionmodules.sampling",
    "sgm.modules.diffusionmodules.sampling_utils", "sgm.modules.encoders", "sgm.modules.encoders.modules",
    "sgm.models", "sgm.models.autoencoder", "sgm.models.diffusion",
)


def install_as_sgm(force: bool = False) -> None:
    """Register this package under the top-level name `sgm` so reference YAML `target:` strings resolve here."""
    if "sgm" in sys.modules and not force and not getattr(sys.modules["sgm"], "__panacea_b200__", False):
        raise RuntimeError("a different `sgm` package is already imported; pass force=True to shadow it")
    for name in _MIRRORED:
        sys.modules[name] = importlib.import_module("panacea_b200." + name)


__panacea_b200__ = True
