"""Dev tool: one warm forward, then ONE forward inside cudaProfilerStart/Stop (use with ncu --profile-from-start off)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from _pkgload import load_package  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--variant", default="base")
ap.add_argument("--task", default="sr")
ap.add_argument("--scale", type=int, default=4)
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--precision", default="bf16")
a = ap.parse_args()
pkg = load_package()
import grl_oracle as orc  # noqa: E402

cfg = pkg.configs.grl_config(a.variant, a.task, a.scale, a.size)
m = pkg.GRL(**cfg)
m.load_state_dict(orc.synth_state_dict(cfg, 0), strict=False)
m = m.cuda().eval()
m.set_precision(a.precision)
x = torch.rand(a.batch, 3, a.size, a.size, device="cuda")
m(x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
m(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
