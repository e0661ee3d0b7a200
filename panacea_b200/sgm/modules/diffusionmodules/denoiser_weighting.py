"""denoiser_weighting.py EpsWeighting (loss weighting; unused at inference, kept so the YAML instantiates)."""


class EpsWeighting:
    def __call__(self, sigma):
        return sigma ** -2.0
