"""sampling_utils.py:7-9,39-40 helpers kept for API parity (the arithmetic lives in pn_cfg_euler_step)."""


class NoDynamicThresholding:
    def __call__(self, uncond, cond, scale):
        return uncond + scale * (cond - uncond)
