"""One eager eps-evaluation of the full-size model (CFG batch, 6 views x 8 frames) inside a cudaProfilerStart/Stop
range, for the per-launch device-time list:
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    pipe = bench.build_pipeline(dev, 3407)
    pipe.wrapper.use_cuda_graph = False
    pipe.wrapper.hint_repeat = 2
    host = bench.synth_inputs_host(3407)
    cc = {"cond_feat": host["hint"].to(dev), "concat": torch.cat([host["concat"]] * 2).to(dev),
          "crossattn": torch.cat([host["uc_txt"], host["c_txt"]]).to(dev)}
    x_in = torch.randn(16, 4, bench.H, bench.VIEWS * bench.W_VIEW, device=dev)
    t = torch.full((16,), 999, dtype=torch.int64, device=dev)
    pipe.wrapper(x_in, t, cc)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    pipe.wrapper(x_in, t, cc)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
