"""denoiser_scaling.py:16-22 EpsScaling as host scalars (the scalings are applied inside fused kernels)."""
import math


class EpsScaling:
    def __call__(self, sigma):
        """Returns (c_skip, c_out, c_in, c_noise) for a python float or a tensor sigma."""
        if isinstance(sigma, (int, float)):
            return 1.0, -float(sigma), 1.0 / math.sqrt(float(sigma) ** 2 + 1.0), float(sigma)
        return sigma * 0 + 1, -sigma, 1 / (sigma ** 2 + 1.0) ** 0.5, sigma.clone()
