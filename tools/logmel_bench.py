"""Times the fused log-mel front end alone through the C ABI (gww_logmel_frontend via logmel_features), CUDA events.
    python tools/logmel_bench.py [--det-windows 296] [--reps 20]        (GWW_LIB=<variant .so> for A/B builds)
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gw_whisper_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--det-windows", type=int, default=296)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
lib = _lib.load()
dev = torch.device("cuda:0")
n = a.det_windows
x = torch.randn(n, 2048, device=dev)
out = torch.zeros(n, 80, 3000, device=dev, dtype=torch.float32)


def run():
    _lib.check(lib.gww_logmel_frontend(x.data_ptr(), n, out.data_ptr(), _lib.stream_ptr()))


for _ in range(3):
    run()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.reps + 1)]
ev[0].record()
for i in range(a.reps):
    run()
    ev[i + 1].record()
torch.cuda.synchronize()
ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(a.reps))
med = ts[len(ts) // 2]
byt = n * (2048 * 4 + 3000 * 80 * 4)
print(json.dumps({"lib": os.path.basename(os.environ.get("GWW_LIB", "default")), "det_windows": n, "ms": round(med, 4),
                  "us_per_det_window_per_sm": round(med * 1e3 * 148 / n, 2), "gbs": round(byt / med / 1e6, 1)}))
