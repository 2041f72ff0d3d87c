"""Times the encoder's GEMM shapes one by one through the C ABI (gww_gemm_bf16), CUDA events on the
launching stream.  Tuning tool: A/B kernel builds with GWW_LIB=..., pair mode with GWW_GEMM_MC=1|2.
    python tools/gemm_bench.py [--model base] [--det-windows 256] [--reps 10] [--only qkv,fc1]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gw_whisper_b200 import _lib  # noqa: E402

DIMS = {"tiny": (384, 1536), "base": (512, 2048), "small": (768, 3072)}

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="base")
ap.add_argument("--det-windows", type=int, default=256)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--only", default="")
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--bn", type=int, default=0, help="force the N tile (128/192/256)")
a = ap.parse_args()
d, ffn = DIMS[a.model]
M = a.det_windows * 1500
lib = _lib.load()
dev = torch.device("cuda:0")
shapes = {  # name: (N, K, epilogue)
    "qkv": (3 * d, d, 0),
    "out_proj": (d, d, 2),
    "fc1": (ffn, d, 1),
    "fc2": (d, ffn, 2),
}
only = [s for s in a.only.split(",") if s]
res = {}
for name, (N, K, epi) in shapes.items():
    if only and name not in only:
        continue
    A = torch.randn(M, K, device=dev).to(_lib.operand_dtype())
    W = (torch.randn(N, K, device=dev) / K ** 0.5).to(_lib.operand_dtype())
    bias = torch.randn(N, device=dev)
    out_f32 = epi in (2, 3)
    Cm = torch.empty(M, N, device=dev, dtype=torch.float32 if out_f32 else _lib.operand_dtype())
    resid = torch.randn(M, N, device=dev) if epi == 2 else None
    bn = a.bn or (256 if N % 256 == 0 else (192 if N % 192 == 0 else 128))

    def run():
        _lib.check(lib.gww_gemm_bf16(A.data_ptr(), W.data_ptr(), Cm.data_ptr(), bias.data_ptr(), _lib.ptr(resid),
                                     None, M, N, K, epi, bn, _lib.stream_ptr()))
    for _ in range(a.warmup):
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
    fl = 2.0 * M * N * K
    byt = M * K * 2 + M * N * (4 if out_f32 else 2) * (2 if epi == 2 else 1)
    res[name] = {"ms": round(med, 4), "min_ms": round(ts[0], 4), "tflops": round(fl / med / 1e9, 1),
                 "hbm_gbs": round(byt / med / 1e6, 1)}
    del A, W, Cm, resid
print(json.dumps({"model": a.model, "M": M, "mc": os.environ.get("GWW_GEMM_MC", "auto"),
                  "lib": os.path.basename(os.environ.get("GWW_LIB", "default")), "gemm": res}))
