"""End-to-end parity of the CUDA path (through the drop-in module API and the C ABI) against
 (a) committed golden outputs of the UNMODIFIED reference (tests/golden/, made by oracle/make_golden.py) and
 (b) the CPU oracle port run on this box on the same seeded inputs.

Two precision modes, two bars:
 * "parity" (ParityOps: split-bf16 operands = fp32-class products on the same tcgen05 kernels, fp32 attention): the
   LITERAL tolerance of BASELINE.json, rtol 1e-3 / atol 1e-4, on >= 99.9 % of the elements of every case, including the
   full-size model at the benchmarked shape [16, 8, 32, 336] (T = 8, CFG b = 2);
 * "bf16" (the benchmarked fast path: bf16 operands, fp32 accumulation / softmax / norm statistics / residual stream):
   the reference's own network under bf16 autocast deviates from its fp32 output by rel-L2 1.7e-2;
   the bound asserted here is ~1.2x what this path measures (rel-L2 <= 9.5e-3, max-abs <= 4.6 % of the output rms; measured
   6.9e-3 .. 7.9e-3 and 2.8 .. 3.8 % with the bf16 token stream inside the transformer blocks), so a regression shows, and the measured rtol/atol pass rate is recorded in gpurun_out/parity.jsonl.
"""
import json
import os
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).resolve().parent / "golden"
BOUNDS = {  # mode -> (rel-L2, max-abs / rms, min fraction inside rtol 1e-3 / atol 1e-4)
    "bf16": (9.5e-3, 0.046, 0.0),
    "parity": (1e-4, 1e-3, 0.999),
    "bf16_loop": (7.5e-3, 0.033, 0.0),    # 25/50 chained bf16 evaluations (measured 6.1e-3 / 4.2e-3, max-abs 2.7 % of rms)
}


def _build(case, use_cuda_graph=False, precision="bf16"):
    from oracle import cases as Cs
    from panacea_b200.pipeline import default_network_config
    from panacea_b200.sgm.modules.diffusionmodules import OpenAIWrapperControlLDM3D
    from panacea_b200.sgm.util import instantiate_from_config
    kw = case.unet_kwargs()
    model = instantiate_from_config(default_network_config(**{k: kw[k] for k in ("model_channels", "num_head_channels", "context_dim", "num_frames")}))
    model.set_precision(precision)
    w = OpenAIWrapperControlLDM3D(model, use_cuda_graph=use_cuda_graph)
    sd = Cs.make_weights(case)
    w.load_state_dict(sd, strict=True)
    return w.cuda(), sd


def _report(name, got, ref, mode="bf16", extra=None):
    d = (got - ref).double()
    rel = (d.norm() / ref.double().norm()).item()
    mx = d.abs().max().item()
    rms = ref.double().pow(2).mean().sqrt().item()
    frac = (d.abs() <= 1e-4 + 1e-3 * ref.double().abs()).double().mean().item()
    rec = {"case": name, "mode": mode, "rel_l2": rel, "max_abs": mx, "ref_rms": rms, "frac_within_rtol1e-3_atol1e-4": frac}
    rec.update(extra or {})
    print("PARITY " + json.dumps(rec))
    out = Path(os.environ.get("PN_PARITY_LOG", "gpurun_out/parity.jsonl"))
    try:
        out.parent.mkdir(parents=True, exist_ok=True)
        with out.open("a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass
    assert torch.isfinite(got).all()
    b_rel, b_max, b_frac = BOUNDS[mode]
    assert rel <= b_rel, f"{name} [{mode}]: rel-L2 {rel:.3e} > {b_rel}"
    assert mx <= b_max * rms, f"{name} [{mode}]: max-abs {mx:.3e} vs rms {rms:.3e}"
    assert frac >= b_frac, f"{name} [{mode}]: only {frac:.5f} of the elements inside rtol 1e-3 / atol 1e-4"
    return rec


@pytest.mark.parametrize("precision", ["bf16", "parity"])
@pytest.mark.parametrize("name", ["small_hd64", "small_hd64_3to2"])
def test_eps_small_vs_reference_golden_and_oracle(name, precision):
    from oracle import cases as Cs, unet_port as P
    case = [c for c in Cs.GOLDEN_CASES if c.name == name][0]
    w, sd = _build(case, precision=precision)
    x, t, c = Cs.make_inputs(case)
    golden = torch.load(GOLDEN / f"eps_{name}.pt")["eps"]
    oracle = P.wrapper_forward(sd, case.net_config(), x, t, c)
    assert (oracle - golden).abs().max().item() < 1e-4, "oracle port drifted from the reference golden"
    cg = {k: v.cuda() for k, v in c.items()}
    eps = w(x.cuda(), t.cuda(), cg).cpu()
    _report(name + ":vs_reference_golden", eps, golden, precision)
    # same again through the CUDA-graph path: must reproduce the eager launch sequence bit for bit
    w.use_cuda_graph = True
    eps_g = w(x.cuda(), t.cuda(), cg).cpu()
    eps_g2 = w(x.cuda(), t.cuda(), cg).cpu()
    assert torch.equal(eps_g, eps) and torch.equal(eps_g2, eps), "CUDA-graph replay differs from eager launches"


def test_module_level_api_matches_fused_path():
    """ControlNet3D.forward / ControlledUNetModel3D.forward (NCHW lists, reference signatures) == wrapper path."""
    from oracle import cases as Cs
    case = [c for c in Cs.GOLDEN_CASES if c.name == "small_hd64"][0]
    w, _ = _build(case)
    x, t, c = Cs.make_inputs(case)
    xg = torch.cat([x, c["concat"]], 1).cuda()
    model = w.diffusion_model
    control = model.controlnet(x=xg, hint=c["cond_feat"].cuda(), timesteps=t.cuda(), context=c["crossattn"].cuda())
    assert len(control) == 13 and control[0].shape == (x.shape[0], 128, case.H, 6 * case.w)
    out = model(xg, timesteps=t.cuda(), context=c["crossattn"].cuda(), control=control)
    assert len(control) == 0
    fused = w(x.cuda(), t.cuda(), {k: v.cuda() for k, v in c.items()})
    assert torch.equal(out, fused)


def test_zero_init_model_predicts_zero_and_cross_view_table():
    """Reference quirks: (i) a freshly constructed model outputs exactly 0 (all zero_module'd tails);
    (ii) perturbing view j changes exactly the views of the asymmetric neighbour table."""
    from oracle import cases as Cs
    from panacea_b200.ops import NativeOps
    from panacea_b200.netplan import CROSS_VIEW_NEIGHBOURS
    from panacea_b200.pipeline import default_network_config
    from panacea_b200.sgm.modules.diffusionmodules import OpenAIWrapperControlLDM3D
    from panacea_b200.sgm.util import instantiate_from_config
    case = [c for c in Cs.GOLDEN_CASES if c.name == "small_hd64"][0]
    kw = case.unet_kwargs()
    model = instantiate_from_config(default_network_config(**{k: kw[k] for k in ("model_channels", "num_head_channels", "context_dim", "num_frames")}))
    w = OpenAIWrapperControlLDM3D(model).cuda()
    x, t, c = Cs.make_inputs(case)
    eps = w(x.cuda(), t.cuda(), {k: v.cuda() for k, v in c.items()})
    assert eps.abs().max().item() == 0.0
    ops = NativeOps()
    Fr, H, wv, heads = 1, 8, 16, 2
    qkv = torch.randn(Fr, H, 6, wv, 3 * 128, device="cuda").to(torch.bfloat16)
    base = ops.attention_view(qkv, heads, True, CROSS_VIEW_NEIGHBOURS)
    expect = {0: {1}, 1: {0, 2}, 2: {1, 3}, 3: {2, 4}, 4: {3, 5}, 5: {0, 4}}   # {v : j in neighbours[v]} — view 5 sees {4} only
    for j in range(6):
        q2 = qkv.clone()
        q2[:, :, j, :, 128:] += 1.0                        # perturb K and V of view j only
        out = ops.attention_view(q2, heads, True, CROSS_VIEW_NEIGHBOURS)
        changed = {v for v in range(6) if not torch.equal(out[:, :, v], base[:, :, v])}
        assert changed == expect[j], (j, changed)


_FULL = {}


def _full_model(num_frames):
    """The full-size (2.24 B parameter) model, built once per frame count and shared by the full-size cases (they all
    use weight seed 0); the precision mode is switched in place (repack on the GPU)."""
    from oracle import cases as Cs
    if num_frames not in _FULL:
        _FULL.clear()                                   # one full-size model resident at a time
        torch.cuda.empty_cache()
        case = [c for c in Cs.GOLDEN_CASES if c.model_channels == 320 and c.num_frames == num_frames][0]
        w, sd = _build(case)
        del sd
        _FULL[num_frames] = w
    return _FULL[num_frames]


@pytest.mark.parametrize("precision", ["bf16", "parity"])
@pytest.mark.parametrize("name", ["full_config1", "full_t1_cond", "full_t8_cfg", "full_2to1_lastframe"])
def test_eps_full_size_vs_reference_golden(name, precision):
    """The full 2.24 B-parameter model against the reference's own output: BASELINE config 1 ([1,8,32,336], T=1, null
    and random conditioning), the BENCHMARKED shape [16,8,32,336] (T=8, CFG b=2; configs 2/3) and config 4 (native 2:1
    32x64 views, T=8, use_last_frame image conditioning). Both precision modes."""
    from oracle import cases as Cs
    case = [c for c in Cs.GOLDEN_CASES if c.name == name][0]
    w = _full_model(case.num_frames)
    w.diffusion_model.set_precision(precision)
    w.invalidate()
    x, t, c = Cs.make_inputs(case)
    golden = torch.load(GOLDEN / f"eps_{name}.pt")["eps"]
    eps = w(x.cuda(), t.cuda(), {k: v.cuda() for k, v in c.items()}).cpu()
    _report(name + ":vs_reference_golden", eps, golden, precision)


def test_two_conditionings_back_to_back_are_not_confused():
    """The step-invariant conditioning cache keys on content, never on addresses: two samples with different BEV hints /
    text through ONE wrapper (tensors freed in between so the allocator recycles their addresses) each match the
    oracle; re-sending equal content in fresh tensors (what the reference's guider does every step) reuses the cache."""
    from oracle import cases as Cs, unet_port as P
    case = [c for c in Cs.GOLDEN_CASES if c.name == "small_hd64"][0]
    w, sd = _build(case)
    eng = w.diffusion_model.engine()
    calls = []
    orig = eng.prepare_condition
    eng.prepare_condition = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    outs = []
    for seed in (1, 2, 1):
        import dataclasses
        cs = dataclasses.replace(case, input_seed=seed)
        x, t, c = Cs.make_inputs(cs)
        cg = {k: v.cuda() for k, v in c.items()}
        eps = w(x.cuda(), t.cuda(), cg).cpu()
        ref = P.wrapper_forward(sd, case.net_config(), x, t, c)
        rel = ((eps - ref).norm() / ref.norm()).item()
        assert rel < 9.5e-3, (seed, rel)
        outs.append(eps)
        n0 = len(calls)
        cg2 = {k: v.clone() for k, v in cg.items()}       # equal content, new tensors: fingerprint hit, no re-preparation
        eps2 = w(x.cuda(), t.cuda(), cg2).cpu()
        assert len(calls) == n0 and torch.equal(eps2, eps)
        del cg, cg2
        torch.cuda.empty_cache()
    assert len(calls) == 3
    assert torch.equal(outs[0], outs[2]) and not torch.equal(outs[0], outs[1])


def _sampler_setup(precision, steps):
    from oracle import cases as Cs
    from oracle.make_golden import sampler_inputs
    from panacea_b200.pipeline import DEFAULT_DENOISER, default_sampler_config
    from panacea_b200.sgm.util import instantiate_from_config
    case = Cs.SAMPLER_CASE
    w, sd = _build(case, precision=precision)
    g = torch.load(GOLDEN / f"sampler_{case.name}_{steps}.pt")
    x, c, uc = sampler_inputs(case, g["use_last_frame"])
    sampler = instantiate_from_config(default_sampler_config(steps, g["scale"]))
    den = instantiate_from_config(DEFAULT_DENOISER)
    return case, w, sd, g, x, c, uc, sampler, den


@pytest.mark.parametrize("precision", ["bf16", "parity"])
@pytest.mark.parametrize("steps", [25, 50])
def test_sampler_loop_vs_reference_golden(steps, precision):
    """The whole EulerEDMSampler + VanillaCFG + DiscreteDenoiser loop (sampling.py:96-133, guiders.py:31-40,
    denoiser.py:22-28) at the reference's 25 steps (with the use_last_frame shared-noise init of DiffusionEngine3D.sample,
    diffusion.py:242-249 = BASELINE config 4) and at BASELINE config 2's 50 steps, against the trajectory the UNMODIFIED
    reference produced on the same seeded inputs. Records the error-vs-step curve; parity mode must meet the literal
    rtol 1e-3 / atol 1e-4 on the final latent."""
    from panacea_b200.pipeline import DenoisingPipeline
    case, w, sd, g, x, c, uc, sampler, den = _sampler_setup(precision, steps)
    pipe = DenoisingPipeline.__new__(DenoisingPipeline)             # assemble from the parts built above
    pipe.model, pipe.wrapper, pipe.denoiser, pipe.sampler = w.diffusion_model, w, den, sampler
    pipe.share_noise_level, pipe.num_frames = g["share_noise_level"], case.num_frames
    traj = []
    sampler.step_callback = lambda i, xx: traj.append(xx.detach().cpu().clone())
    cg = {k: v.cuda() for k, v in c.items()}
    ucg = {k: v.cuda() for k, v in uc.items()}
    out = pipe.sample(cg, ucg, x.cuda(), share_noise=g["use_last_frame"]).cpu()
    assert sampler.last_timestep_indices == g["timestep_indices"]
    ref_steps = g["x_steps"]                                       # x at the START of step i (i = 0 .. steps-1)
    curve = []
    for i in range(1, steps):
        d = (traj[i - 1] - ref_steps[i]).double()
        curve.append(float(d.norm() / ref_steps[i].double().norm()))
    d = (out - g["x_final"]).double()
    curve.append(float(d.norm() / g["x_final"].double().norm()))
    # The seeded random-weight network is no denoiser: the latent stays at the sigma_0 scale (rms 16.3) instead of ending
    # at unit scale like a trained model's. rtol/atol are therefore applied to the latent normalised by the reference's
    # rms (atol 1e-4 is a unit-scale bound); the raw-scale pass rate is recorded next to it.
    rms = g["x_final"].double().pow(2).mean().sqrt().item()
    raw_frac = (d.abs() <= 1e-4 + 1e-3 * g["x_final"].double().abs()).double().mean().item()
    rec = _report(f"sampler{steps}{'_last_frame' if g['use_last_frame'] else ''}:x_final_vs_reference", out / rms, g["x_final"] / rms,
                  "parity" if precision == "parity" else "bf16_loop",
                  {"rel_l2_after_step": [round(v, 7) for v in curve], "latent_rms": rms, "frac_within_tol_at_raw_scale": raw_frac})
    assert rec is not None


def test_sampler_accepts_a_plain_callable_like_the_reference():
    """diffusion.py:251-254 passes a lambda (denoiser closing over the model) to the sampler: the mirror sampler must take
    any callable; the result equals the fused BoundDenoiser path up to the rounding of one extra fp32 pass per step."""
    from panacea_b200.sgm.modules.diffusionmodules.sampling import BoundDenoiser
    case, w, sd, g, x, c, uc, sampler, den = _sampler_setup("bf16", 25)
    cg = {k: v.cuda() for k, v in c.items()}
    ucg = {k: v.cuda() for k, v in uc.items()}
    steps = 4
    fused = sampler(BoundDenoiser(den, w), x.cuda(), cg, ucg, num_steps=steps).cpu()
    plain = sampler(lambda xx, sigma, cc: den(w, xx, sigma, cc), x.cuda(), cg, ucg, num_steps=steps).cpu()
    rel = ((fused - plain).norm() / fused.norm()).item()
    assert rel < 5e-3, rel      # a 1-ulp difference in x flips bf16 roundings inside the network
