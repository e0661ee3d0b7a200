"""Drop-in for `sgm.modules.diffusionmodules.wrappers.OpenAIWrapperControlLDM3D` (reference wrappers.py:37-70):
`forward(x, t, c) -> eps` with c = {concat, cond_feat, crossattn}. This is the operator boundary the denoiser
calls once per Euler step (denoiser.py:28)."""
from __future__ import annotations

import torch
import torch.nn as nn

OPENAIUNETWRAPPERCONTROLLDM3D = "sgm.modules.diffusionmodules.wrappers.OpenAIWrapperControlLDM3D"


class IdentityWrapper(nn.Module):
    def __init__(self, diffusion_model, compile_model: bool = False):
        super().__init__()
        if compile_model:
            raise NotImplementedError("compile_model=True: torch.compile is not part of this implementation")
        self.diffusion_model = diffusion_model

    def forward(self, *args, **kwargs):
        return self.diffusion_model(*args, **kwargs)


def _cond_key(t: torch.Tensor):
    return (t.data_ptr(), t._version, tuple(t.shape), t.dtype)


class OpenAIWrapperControlLDM3D(IdentityWrapper):
    """ControlNet -> UNet on the channel-concatenated latent. Step-invariant work (hint stem, text K/V) is cached and
    reused while `c["cond_feat"]` / `c["crossattn"]` are the same tensors; the per-step network runs as one CUDA
    graph when `use_cuda_graph` is set (static shapes, no allocation inside the graph's replay)."""

    def __init__(self, diffusion_model, compile_model: bool = False, use_cuda_graph: bool = False, hint_repeat: int = 1):
        super().__init__(diffusion_model, compile_model)
        self.use_cuda_graph = use_cuda_graph
        self.hint_repeat = hint_repeat        # 2 when cond_feat holds the hint once for both CFG halves
        self._cond_id = None
        self._graph = None
        self._graph_sig = None
        self._static = None

    def invalidate(self, drop_graph: bool = True) -> None:
        """Forget the prepared conditioning (next call re-runs the hint stem and the text K/V projections). The captured
        graph only depends on shapes and on buffer addresses that survive a new conditioning, so a serving loop passes
        drop_graph=False between samples and keeps replaying it."""
        self._cond_id = None
        if drop_graph:
            self._graph = None

    @torch.no_grad()
    def prepare(self, c: dict) -> None:
        eng = self.diffusion_model.engine()
        cid = (_cond_key(c["cond_feat"]), _cond_key(c["crossattn"]), id(eng.wu), self.hint_repeat)
        if cid != self._cond_id:
            eng.prepare_condition(c["cond_feat"].float(), c["crossattn"].float(), hint_repeat=self.hint_repeat)
            self._cond_id = cid       # condition buffers keep their addresses, so a captured graph stays valid

    @torch.no_grad()
    def forward(self, x: torch.Tensor, t: torch.Tensor, c: dict, **kwargs) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("panacea_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        self.prepare(c)
        eng = self.diffusion_model.engine()
        x = x.float().contiguous()
        t = t.to(torch.int64).contiguous()
        concat = c.get("concat", None)
        concat = None if concat is None else concat.float().contiguous()
        if not self.use_cuda_graph:
            return eng.eps(x, concat, t)
        sig = (tuple(x.shape), None if concat is None else tuple(concat.shape), tuple(c["cond_feat"].shape),
               tuple(c["crossattn"].shape), id(eng.wu), id(eng.cond["guided"]))
        if self._graph is None or self._graph_sig != sig:
            self._capture(eng, x, concat, t, sig)
        sx, sc, st, so = self._static
        sx.copy_(x)
        st.copy_(t)
        if sc is not None:
            sc.copy_(concat)
        self._graph.replay()
        return so.clone()

    def _capture(self, eng, x, concat, t, sig):
        sx, st = x.clone(), t.clone()
        sc = None if concat is None else concat.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            eng.eps(sx, sc, st)          # warm-up: function attributes, TMA-map cache, allocator pools
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            so = eng.eps(sx, sc, st)
        self._graph, self._graph_sig, self._static = g, sig, (sx, sc, st, so)
