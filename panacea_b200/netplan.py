"""Network plan: configuration, block topology and parameter specification of the denoising networks.

The product's own description of what `configs/inference_nuscenes.yaml:30-71` builds in the reference
(`UNetModel3D.__init__`, openaimodel.py:804-1261; `ControlNet3D.__init__`, controlmodel.py:26-84). It yields
(a) the ordered list of stages the engine executes and (b) the exact state-dict keys/shapes, so checkpoints
written for the reference load unchanged.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Iterable

# Reference quirk kept on purpose (attention.py:545-559): the cross-view ring is asymmetric — view 0 sees {5,1},
# views 1..4 see {v-1,v+1}, view 5 sees {4} only. Exposed as data so a symmetric ring is a one-line change.
CROSS_VIEW_NEIGHBOURS = ((5, 1), (0, 2), (1, 3), (2, 4), (3, 5), (4,))
HINT_CHANNELS = (16, 16, 32, 32, 96, 96, 256)   # controlmodel.py:43-59 (last conv -> model_channels)
HINT_STRIDES = (1, 1, 2, 1, 2, 1, 2, 1)


@dataclass
class NetConfig:
    in_channels: int = 8
    out_channels: int = 4
    model_channels: int = 320
    attention_resolutions: tuple = (4, 2, 1)
    num_res_blocks: int = 2
    channel_mult: tuple = (1, 2, 4, 4)
    num_head_channels: int = 64
    context_dim: int = 1024
    num_frames: int = 8
    hint_channels: int = 19
    control_scales: float = 1.0
    num_views: int = 6

    @property
    def emb_channels(self) -> int:
        return 4 * self.model_channels


UNSUPPORTED_KW = {
    "dims": 2, "num_classes": None, "num_heads": -1, "num_heads_upsample": -1, "use_scale_shift_norm": False,
    "resblock_updown": False, "use_new_attention_order": False, "n_embed": None, "disable_self_attentions": None,
    "num_attention_blocks": None, "disable_middle_self_attn": False, "adm_in_channels": None,
    "transformer_depth_middle": None, "dropout": 0, "conv_resample": True,
}


def config_from_kwargs(kw: dict, hint_channels: int = 19, control_scales: float = 1.0) -> NetConfig:
    """Validates reference constructor kwargs; only the configuration family the YAML uses is implemented."""
    kw = dict(kw)
    for k, default in UNSUPPORTED_KW.items():
        if k in kw and kw[k] not in (default, None if default is None else default):
            raise NotImplementedError(f"panacea_b200: {k}={kw[k]!r} is not supported (reference default {default!r} only)")
    if not kw.get("use_spatial_transformer", False) or not kw.get("use_linear_in_transformer", False):
        raise NotImplementedError("panacea_b200 implements use_spatial_transformer=True, use_linear_in_transformer=True")
    if kw.get("spatial_only_attn_type") != "intra-view" or not kw.get("insert_crossview", False):
        raise NotImplementedError("panacea_b200 implements spatial_only_attn_type='intra-view' with insert_crossview=True")
    if kw.get("legacy", True):
        raise NotImplementedError("legacy=True head sizing is not supported")
    td = kw.get("transformer_depth", 1)
    if (td if isinstance(td, int) else max(td)) != 1:
        raise NotImplementedError("transformer_depth must be 1")
    nrb = kw["num_res_blocks"]
    if not isinstance(nrb, int):
        if len(set(nrb)) != 1:
            raise NotImplementedError("per-level num_res_blocks is not supported")
        nrb = nrb[0]
    if kw["num_head_channels"] in (-1, None):
        raise NotImplementedError("num_head_channels must be set")
    return NetConfig(
        in_channels=kw["in_channels"], out_channels=kw.get("out_channels", kw["in_channels"]),
        model_channels=kw["model_channels"], attention_resolutions=tuple(kw["attention_resolutions"]),
        num_res_blocks=nrb, channel_mult=tuple(kw["channel_mult"]), num_head_channels=kw["num_head_channels"],
        context_dim=kw["context_dim"], num_frames=kw.get("num_frames", 4), hint_channels=hint_channels,
        control_scales=control_scales)


@dataclass
class Stage:
    kind: str        # "stem" | "res" | "stt" | "down" | "up"
    key: str
    cin: int
    cout: int
    heads: int = 0


@dataclass
class Plan:
    encoder: list = field(default_factory=list)   # list[list[Stage]] — one entry per input block
    middle: list = field(default_factory=list)
    decoder: list = field(default_factory=list)
    skip_channels: list = field(default_factory=list)

    def stages(self) -> Iterable[Stage]:
        for blk in self.encoder:
            yield from blk
        yield from self.middle
        for blk in self.decoder:
            yield from blk


def make_plan(cfg: NetConfig, decoder: bool) -> Plan:
    mc = cfg.model_channels
    plan = Plan()
    plan.encoder.append([Stage("stem", "input_blocks.0.0", cfg.in_channels, mc)])
    widths = [mc]
    ch, ds, n = mc, 1, 1
    last = len(cfg.channel_mult) - 1
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            blk = [Stage("res", f"input_blocks.{n}.0", ch, mult * mc)]
            ch = mult * mc
            if ds in cfg.attention_resolutions:
                blk.append(Stage("stt", f"input_blocks.{n}.1", ch, ch, ch // cfg.num_head_channels))
            plan.encoder.append(blk)
            widths.append(ch)
            n += 1
        if level != last:
            plan.encoder.append([Stage("down", f"input_blocks.{n}.0", ch, ch)])
            widths.append(ch)
            n += 1
            ds *= 2
    plan.skip_channels = list(widths)
    plan.middle = [Stage("res", "middle_block.0", ch, ch),
                   Stage("stt", "middle_block.1", ch, ch, ch // cfg.num_head_channels),
                   Stage("res", "middle_block.2", ch, ch)]
    if decoder:
        n = 0
        for level in range(last, -1, -1):
            mult = cfg.channel_mult[level]
            for i in range(cfg.num_res_blocks + 1):
                blk = [Stage("res", f"output_blocks.{n}.0", ch + widths.pop(), mc * mult)]
                ch = mc * mult
                if ds in cfg.attention_resolutions:
                    blk.append(Stage("stt", f"output_blocks.{n}.1", ch, ch, ch // cfg.num_head_channels))
                if level and i == cfg.num_res_blocks:
                    blk.append(Stage("up", f"output_blocks.{n}.{len(blk)}", ch, ch))
                    ds //= 2
                plan.decoder.append(blk)
                n += 1
    return plan


# ------------------------------------------------------------------------------------------------ parameter spec
def _norm(spec, key, c):
    spec[key + ".weight"] = (c,)
    spec[key + ".bias"] = (c,)


def _affine(spec, key, cout, *rest):
    spec[key + ".weight"] = (cout, *rest)
    spec[key + ".bias"] = (cout,)


def _res_params(spec, k, cin, cout, emb):
    _norm(spec, k + ".in_layers.0", cin)
    _affine(spec, k + ".in_layers.2", cout, cin, 3, 3)
    _norm(spec, k + ".in_layers_temporal.0", cout)
    _affine(spec, k + ".in_layers_temporal.2", cout, cout, 3)
    _affine(spec, k + ".emb_layers.1", cout, emb)
    _norm(spec, k + ".out_layers.0", cout)
    _affine(spec, k + ".out_layers.3", cout, cout, 3, 3)
    _norm(spec, k + ".out_layers_temporal.0", cout)
    _affine(spec, k + ".out_layers_temporal.3", cout, cout, 3)
    if cin != cout:
        _affine(spec, k + ".skip_connection", cout, cin, 1, 1)


STT_BRANCHES = ("", "_crossview", "_temporal")   # execution order: intra-view, cross-view, temporal


def _stt_params(spec, k, c, ctx):
    for br in STT_BRANCHES:
        _norm(spec, f"{k}.norm{br}", c)
        _affine(spec, f"{k}.proj_in{br}", c, c)
        _affine(spec, f"{k}.proj_out{br}", c, c)
        t = f"{k}.transformer_blocks{br}.0"
        for attn, kd in (("attn1", c), ("attn2", ctx)):
            spec[f"{t}.{attn}.to_q.weight"] = (c, c)
            spec[f"{t}.{attn}.to_k.weight"] = (c, kd)
            spec[f"{t}.{attn}.to_v.weight"] = (c, kd)
            _affine(spec, f"{t}.{attn}.to_out.0", c, c)
        _affine(spec, f"{t}.ff.net.0.proj", 8 * c, c)
        _affine(spec, f"{t}.ff.net.2", c, 4 * c)
        for nm in ("norm1", "norm2", "norm3"):
            _norm(spec, f"{t}.{nm}", c)


def _trunk_params(spec, plan: Plan, cfg: NetConfig):
    emb = cfg.emb_channels
    _affine(spec, "time_embed.0", emb, cfg.model_channels)
    _affine(spec, "time_embed.2", emb, emb)
    for st in plan.stages():
        if st.kind == "stem":
            _affine(spec, st.key, st.cout, st.cin, 3, 3)
        elif st.kind == "res":
            _res_params(spec, st.key, st.cin, st.cout, emb)
        elif st.kind == "stt":
            _stt_params(spec, st.key, st.cin, cfg.context_dim)
        elif st.kind == "down":
            _affine(spec, st.key + ".op", st.cout, st.cin, 3, 3)
        elif st.kind == "up":
            _affine(spec, st.key + ".conv", st.cout, st.cin, 3, 3)


def unet_param_spec(cfg: NetConfig) -> dict:
    """Keys/shapes of ControlledUNetModel3D's OWN parameters (without the nested `controlnet.` subtree)."""
    spec: dict = {}
    _trunk_params(spec, make_plan(cfg, decoder=True), cfg)
    _norm(spec, "out.0", cfg.model_channels)
    _affine(spec, "out.2", cfg.out_channels, cfg.model_channels, 3, 3)
    return spec


def controlnet_param_spec(cfg: NetConfig) -> dict:
    spec: dict = {}
    plan = make_plan(cfg, decoder=False)
    _trunk_params(spec, plan, cfg)
    chans = (cfg.hint_channels, *HINT_CHANNELS, cfg.model_channels)
    for i in range(len(chans) - 1):
        _affine(spec, f"input_hint_block.{2 * i}", chans[i + 1], chans[i], 3, 3)
    for i, c in enumerate(plan.skip_channels):
        _affine(spec, f"zero_convs.{i}.0", c, c, 1, 1)
    c = plan.skip_channels[-1]
    _affine(spec, "middle_block_out.0", c, c, 1, 1)
    return spec


# Parameters the reference zero-initialises (openaimodel.py:418,454-476,1251; attention.py:1040-1059;
# controlmodel.py:58,82-84). Used only to reproduce the reference's default init.
def is_zero_init(key: str) -> bool:
    tails = (".in_layers_temporal.2.", ".out_layers.3.", ".out_layers_temporal.3.", "proj_out.", "proj_out_temporal.",
             "proj_out_crossview.", "zero_convs.", "middle_block_out.", "input_hint_block.14.")
    if key.startswith("out.2."):
        return True
    return any(t in key for t in tails)
