"""Inference entry point with the control flow of the reference's inference.py (main, :230-317): one process per GPU,
`DistributedSampler(shuffle=False)` + batch size 1 (one 6-view x 8-frame sequence per step and rank), seed = rank + 3407,
model built from `configs/inference_nuscenes.yaml`-style YAML via instantiate_from_config, `log_images` per batch, frame
writers. SURVEY.md section 8f rows N1 (engine / conditioner glue) and N3 (writers, gather -> rank-0 writer).

What is NOT here, and why: the nuScenes dataset + BEV rasteriser (row N4, needs nuScenes and mmdet3d) is replaced by
`SyntheticBEVDataset` with the same batch contract; the CLIP text tower is a deterministic stand-in (BASELINE.json
configs[3]); the VAE (encoder and decoder) is native and random-init unless a checkpoint provides `first_stage_model.*`.
With the real modules importable, `--dataset module:Class` and the YAML targets swap them in.

  torchrun --nproc-per-node 8 -m panacea_b200.inference --base configs.yaml --name run1 --inferdir out --gather
"""
from __future__ import annotations

import argparse
import importlib
import os
import time

import torch
import torch.distributed as dist
import yaml
from torch.utils.data import DataLoader, Dataset
from torch.utils.data.distributed import DistributedSampler

from . import dist_utils as D
from . import frame_io as IO
from .sgm.util import instantiate_from_config


class SyntheticBEVDataset(Dataset):
    """Batch contract of sgm/data/nuscenes_video/nuscenes_datasets_video.py:495-570 (`MyDataset.__getitem__`) with
    synthetic content: `jpg` target frames [T,3,H,6w] in [-1,1], `cond_img` the 19-channel BEV control maps [T,19,H,6w]
    in [0,1], `final_cond_zero` the image condition (zeros except the last — or first — frame, :559-566), `txt`,
    `filenames` (per frame, per camera)."""

    def __init__(self, num_sequences=2, num_frames=8, image_hw=(256, 512), use_last_frame=True, seed=0):
        self.n, self.T, (self.h, self.w), self.use_last_frame, self.seed = num_sequences, num_frames, image_hw, use_last_frame, seed

    def __len__(self):
        return self.n

    def __getitem__(self, idx):
        g = torch.Generator().manual_seed(self.seed * 100003 + idx)
        W = 6 * self.w
        target = torch.rand(self.T, 3, self.h, W, generator=g) * 2.0 - 1.0
        cond = torch.zeros_like(target)
        k = -1 if self.use_last_frame else 0
        cond[k] = target[k]
        scene = f"n015-2018-07-24-11-22-45+0800__seq{idx:04d}"
        names = [[f"samples/{cam}/{scene}__{cam}__{1532402927 + 50 * f:d}.jpg" for cam in
                  sorted(IO.VIEW_ID, key=IO.VIEW_ID.get)] for f in range(self.T)]
        return {"jpg": target, "cond_img": torch.rand(self.T, 19, self.h, W, generator=g), "final_cond_zero": cond,
                "txt": "a driving scene, six surround-view cameras", "filenames": names}


def load_config(paths, overrides=()):
    cfg = {}
    for p in paths:
        with open(p) as f:
            new = yaml.safe_load(f)
        cfg = _merge(cfg, new)
    for ov in overrides:                                       # key.sub.key=value (OmegaConf dotlist style)
        k, v = ov.split("=", 1)
        node = cfg
        parts = k.split(".")
        for part in parts[:-1]:
            node = node.setdefault(part, {})
        node[parts[-1]] = yaml.safe_load(v)
    return cfg


def _merge(a, b):
    if isinstance(a, dict) and isinstance(b, dict):
        out = dict(a)
        for k, v in b.items():
            out[k] = _merge(a[k], v) if k in a else v
        return out
    return b


def model_load_ckpt(model, path):
    """inference.py:198-228: engine checkpoints (.ckpt, DeepSpeed prefix `_forward_module.` stripped) or safetensors;
    loaded non-strictly like the reference."""
    if path.endswith("ckpt"):
        sd = torch.load(path, map_location="cpu")
        sd = sd.get("state_dict", sd.get("module", sd))
        sd = {k.replace("_forward_module.", ""): v for k, v in sd.items()}
    elif path.endswith("safetensors"):
        from safetensors.torch import load_file
        sd = load_file(path)
    else:
        raise NotImplementedError(path)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    print(f"Restored from {path} with {len(missing)} missing and {len(unexpected)} unexpected keys")
    return model


def get_parser():
    p = argparse.ArgumentParser()
    p.add_argument("-n", "--name", type=str, default="")
    p.add_argument("-b", "--base", nargs="*", default=[], help="YAML configs, merged left to right")
    p.add_argument("--inferdir", type=str, default="inferences")
    p.add_argument("--ckptpath", type=str, default=None)
    p.add_argument("--split", type=str, default="val")
    p.add_argument("--use_last_frame", type=lambda v: str(v).lower() in ("1", "true", "yes", "y", "t"), default=True)
    p.add_argument("-s", "--seed", type=int, default=D.BASE_SEED)
    p.add_argument("--bs", type=int, default=1)
    p.add_argument("--dataset", type=str, default=None, help="module:Class of a dataset with the MyDataset batch contract")
    p.add_argument("--num_sequences", type=int, default=2)
    p.add_argument("--image_hw", type=int, nargs=2, default=(256, 512), help="per-view image size of the synthetic dataset")
    p.add_argument("--gather", action="store_true", help="gather decoded frames on rank 0 and let rank 0 write them")
    p.add_argument("--randomize_zero_init", action="store_true", help="re-draw the reference's zero-initialised tails (no checkpoint)")
    return p


def main(argv=None):
    opt, unknown = get_parser().parse_known_args(argv)
    if not opt.name:
        raise ValueError("You must specify the experiment name!!")
    assert opt.bs == 1, "the reference runs batch size 1 (one sequence per rank and step)"
    inferdir = os.path.join(opt.inferdir, opt.name)
    config = load_config(opt.base, unknown)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
        rank = dist.get_rank()
    else:
        rank, local = 0, 0
        torch.cuda.set_device(0)
    seed = rank + opt.seed                                     # inference.py:250
    torch.manual_seed(seed)
    device = torch.device("cuda", local)

    if opt.dataset:
        mod, cls = opt.dataset.split(":")
        dataset = getattr(importlib.import_module(mod), cls)(split=opt.split, use_last_frame=opt.use_last_frame)
    else:
        T = config["model"]["params"]["network_config"]["params"].get("num_frames", 8)
        dataset = SyntheticBEVDataset(opt.num_sequences, T, tuple(opt.image_hw), opt.use_last_frame, seed=opt.seed)
    sampler = DistributedSampler(dataset, num_replicas=world, rank=rank, shuffle=False)
    loader = DataLoader(dataset, batch_size=opt.bs, sampler=sampler)

    model = instantiate_from_config(config["model"])
    if opt.ckptpath is not None:
        model = model_load_ckpt(model, opt.ckptpath)
    elif opt.randomize_zero_init:
        model.model.diffusion_model.randomize_zero_init(seed=opt.seed)
        model.model.diffusion_model.controlnet.randomize_zero_init(seed=opt.seed + 1)
    model.to(device).eval()

    all_time, written = 0.0, []
    for idx, batch in enumerate(loader):
        start = time.time()
        for key in batch:
            if key not in ("txt", "filenames"):
                batch[key] = batch[key].to(device)
        with torch.no_grad():
            outs = model.log_images(batch)
        filenames = batch["filenames"]
        samples = outs["samples"]
        if opt.gather and world > 1:                           # BASELINE.json configs[2]: NCCL gather of decoded frames
            gathered = D.gather_on_rank0(samples.contiguous())
            names = [None] * world
            dist.all_gather_object(names, filenames)
            if rank == 0:
                for r in range(world):
                    written += IO.logs_frames(gathered[r], os.path.join(inferdir, "fake"), names[r])
        else:
            written += IO.logs_frames(samples, os.path.join(inferdir, "fake"), filenames)
        written += IO.logs_all_images(outs, os.path.join(inferdir, "allimages"), filenames)
        written += IO.logs_all_gifs(outs, os.path.join(inferdir, "gifs"), filenames, num_frames=samples.shape[0])
        all_time += time.time() - start
        if rank == 0:
            print(f"idx {idx}: time per iter {time.time() - start:.2f}s, avg {all_time / (idx + 1):.2f}s", flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return written


if __name__ == "__main__":
    main()
