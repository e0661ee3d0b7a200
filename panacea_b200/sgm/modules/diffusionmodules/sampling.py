"""sampling.py:24-133,214-218 EulerEDMSampler (s_churn = 0): deterministic Euler steps in sigma space with
classifier-free guidance — "DDIM (eta=0)" in the reference config (configs/inference_nuscenes.yaml:115-126).

Per step the device executes: one eps evaluation of the CFG-doubled batch (ControlNet + UNet, optionally one CUDA
graph replay) and ONE fused kernel (pn_cfg_euler_step) that applies the denoiser scalings, the guidance
combination, the Euler update and the next step's input scaling + batch doubling. The reference's per-step
torch.cat / dict rebuild (guiders.py:31-40) happens once per sample here."""
from __future__ import annotations

import math

import torch

from ...util import default, instantiate_from_config

DEFAULT_GUIDER = {"target": "sgm.modules.diffusionmodules.guiders.IdentityGuider"}


class BoundDenoiser:
    """What `DiffusionEngine3D.sample` passes to the sampler in the reference is a lambda closing over
    (denoiser, model) (diffusion.py:251-254); this object is the same callable with the two parts visible, so the
    sampler can fuse the denoiser scalings, the guidance and the Euler update into one kernel per step."""

    def __init__(self, denoiser, network):
        self.denoiser = denoiser
        self.network = network

    def __call__(self, x, sigma, cond):
        return self.denoiser(self.network, x, sigma, cond)


class EulerEDMSampler:
    def __init__(self, discretization_config, num_steps=None, guider_config=None, verbose=False, device="cuda",
                 s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0):
        if s_churn != 0.0:
            raise NotImplementedError("s_churn > 0 (stochastic churn) is not used by the reference config")
        self.num_steps = num_steps
        self.discretization = instantiate_from_config(discretization_config)
        self.guider = instantiate_from_config(default(guider_config, DEFAULT_GUIDER))
        self.verbose = verbose
        self.device = device
        self.last_timestep_indices = []
        self.step_callback = None           # optional: called as step_callback(i, x) after every Euler step (tests)

    def sigmas(self, num_steps=None):
        return self.discretization(self.num_steps if num_steps is None else num_steps, device="cpu")

    @torch.no_grad()
    def __call__(self, denoiser, x, cond, uc=None, num_steps=None):
        """x [N,4,H,W] initial noise (unit variance); cond / uc dicts as produced by the conditioner. `denoiser` is any
        callable (x, sigma, cond) -> denoised, like the reference's lambda (diffusion.py:251-254); a `BoundDenoiser`
        takes the fused path. Returns the final latent, like the reference."""
        from ....ops import NativeOps
        ops = NativeOps()
        uc = default(uc, cond)
        sig = [float(s) for s in self.sigmas(num_steps)]
        cfg = hasattr(self.guider, "scale") and not type(self.guider).__name__.startswith("Identity")
        if not cfg:
            raise NotImplementedError("this sampler implements VanillaCFG (the reference config)")
        n = x.shape[0]
        # prepare_sampling_loop (sampling.py:44-55): x *= sqrt(1 + sigma_0^2)
        x = ops.scale_dup(x.float().contiguous(), math.sqrt(1.0 + sig[0] ** 2.0), 1)
        if not isinstance(denoiser, BoundDenoiser):
            return self._generic_loop(ops, denoiser, x, cond, uc, sig)
        den, net = denoiser.denoiser, denoiser.network
        cc = self.guider.prepare_cond(cond, uc)                   # once per sample
        if hasattr(net, "prepare"):
            net.prepare(cc)                                       # step-invariant conditioning work, once per sample
        scal = [den.step_scalars(s) for s in sig[:-1]]
        self.last_timestep_indices = [s[0] for s in scal]
        t_all = torch.tensor([[s[0]] * (2 * n) for s in scal], dtype=torch.int64, device=x.device)
        x_in = ops.scale_dup(x, scal[0][2], 2)                    # input * c_in, CFG batch doubling
        for i in range(len(sig) - 1):
            eps = net(x_in, t_all[i], cc, return_static=True) if hasattr(net, "static_io") else net(x_in, t_all[i], cc)
            c_in_next = scal[i + 1][2] if i + 1 < len(scal) else 0.0
            ops.cfg_euler_step(x, eps, x_in, sig[i], sig[i + 1], self.guider.scale, c_in_next, sigma_q=scal[i][1])
            if self.step_callback is not None:
                self.step_callback(i, x)
        return x

    def _generic_loop(self, ops, denoiser, x, cond, uc, sig):
        """The reference's call contract (sampling.py:85-133): an opaque `denoiser(x2, sigma2, cond2) -> denoised2` per
        step with the guider's doubled inputs; guidance combination + Euler update in one kernel (net_is_denoised)."""
        n = x.shape[0]
        self.last_timestep_indices = []
        for i in range(len(sig) - 1):
            s_in = torch.full((n,), sig[i], dtype=torch.float32, device=x.device)
            x2, s2, c2 = self.guider.prepare_inputs(x, s_in, cond, uc)
            den2 = denoiser(x2, s2, c2).float().contiguous()
            ops.cfg_euler_step(x, den2, None, sig[i], sig[i + 1], self.guider.scale, 0.0, net_is_denoised=True)
            if self.step_callback is not None:
                self.step_callback(i, x)
        return x
