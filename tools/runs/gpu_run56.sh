#!/bin/bash
# run 56: Q-Adapter convolutions with packed FFMA2: parity (bit-identical maths) and the MLGWSC workload
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_qfront_gpu.py -m gpu -q -x -s -p no:cacheprovider > gpurun_out/q_qfront.log 2>&1; echo "qfront tests rc $?"; grep -E "err|passed|failed" gpurun_out/q_qfront.log | tail -8 | cut -c1-200
timeout 300 python bench.py --workload mlgwsc --model tiny --batch 1024 --steps 3 --warmup 3 > gpurun_out/bench_mlgwsc.log 2> gpurun_out/bench_mlgwsc.err; echo "mlgwsc rc $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_mlgwsc.log").read().strip().splitlines()[-1])
print("value",round(d["value"],1), "ms",round(d["ms_per_step"],1))
print({k:round(v["ms_per_step"],2) for k,v in d["kernels"].items()})
PY
