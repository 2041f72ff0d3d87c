"""One chunk (256 det-windows) of the bench workload, twice: the command profiled under ncu.
    python tools/profile_step.py [--model base] [--windows 128]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import build_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="base")
ap.add_argument("--windows", type=int, default=128)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
model, _ = build_model(a.model, 2 * a.windows)
g = torch.Generator().manual_seed(1234)
strain = torch.randn(a.windows, 2, 2048, generator=g).cuda()
for _ in range(a.reps):
    out = model.forward_strain(strain)
torch.cuda.synchronize()
print("ok", float(out.mean()))
