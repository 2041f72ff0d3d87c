"""Times the QScan front end alone (qscan_tiles_kernel + plane arg-max + qscan_interp_kernel) through QScanB200, CUDA events.
    python tools/qscan_bench.py [--det-windows 256] [--reps 10]        (GWW_LIB=<variant .so> for A/B builds)
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gw_whisper_b200.qfrontend import QScanB200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--det-windows", type=int, default=256)
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
q = QScanB200(1.0, 2048, (512, 512), (4, 128))
x = torch.randn(a.det_windows, 2048, device="cuda")
for _ in range(3):
    y = q(x)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.reps + 1)]
ev[0].record()
for i in range(a.reps):
    y = q(x)
    ev[i + 1].record()
torch.cuda.synchronize()
ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(a.reps))
print(json.dumps({"lib": os.path.basename(os.environ.get("GWW_LIB", "default")), "det_windows": a.det_windows,
                  "ms": round(ts[len(ts) // 2], 4), "min_ms": round(ts[0], 4)}))
