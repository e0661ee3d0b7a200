from .encoders.modules import GeneralConditioner  # noqa: F401  (configs/inference_nuscenes.yaml:73 targets sgm.modules.GeneralConditioner)
