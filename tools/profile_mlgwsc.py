"""One 256-window batch of the MLGWSC-1 model (QScan + Q-Adapter + whisper-tiny), twice: the command profiled under
`ncu --set full` for the front-end kernels.
    python tools/profile_mlgwsc.py [--windows 256]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import HOP, build_mlgwsc_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--windows", type=int, default=256)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
model, _ = build_mlgwsc_model("tiny", 256)
g = torch.Generator().manual_seed(1234)
seg = torch.randn(2, 2048 + HOP * (a.windows - 1), generator=g).cuda()
for _ in range(a.reps):
    scores, idx, sc = model.stream_search(seg, HOP, a.windows, 0.5)
torch.cuda.synchronize()
print("ok", float(scores.mean()))
