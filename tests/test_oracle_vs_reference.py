"""Live comparison of the oracle port with the reference itself — only where /root/reference exists (the build
container). On the GPU box this module is skipped; the committed goldens (test_oracle_golden.py) stand in."""
import pytest
import torch

from oracle import cases as Cs
from oracle import ref_loader as R
from oracle import unet_port as P

pytestmark = pytest.mark.skipif(not R.reference_available(), reason="/root/reference not present")


def test_reference_state_dict_keys_and_port_output():
    case = Cs.GOLDEN_CASES[1]                      # non-2:1 views -> exercises the view-height shim
    model = R.build_reference_model(case.unet_kwargs())
    sd_ref = model.state_dict()
    spec = P.state_spec(case.net_config())
    assert {k: tuple(v.shape) for k, v in sd_ref.items()} == spec
    sd = Cs.make_weights(case)
    model.load_state_dict(sd, strict=True)
    x, t, c = Cs.make_inputs(case)
    with torch.no_grad(), R.view_height_shim(case.H, case.w):
        ref = model(x, t, dict(c))
    out = P.wrapper_forward(sd, case.net_config(), x, t, c)
    assert (out - ref).abs().max().item() < 2e-5


def test_fresh_reference_model_outputs_zero():
    """zero_module'd tails make a fresh model predict exactly 0."""
    case = Cs.GOLDEN_CASES[0]
    model = R.build_reference_model(case.unet_kwargs())
    x, t, c = Cs.make_inputs(case)
    with torch.no_grad():
        assert model(x, t, dict(c)).abs().max().item() == 0.0
