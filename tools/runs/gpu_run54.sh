#!/bin/bash
# run 54: final validation of the round: pytest -m gpu exactly as the driver runs it, smoke(), both bench arms
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1 ) 2> gpurun_out/pytest_gpu.time; echo "pytest -m gpu rc $?"; tail -n 3 gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/pytest_gpu.time
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "build+smoke rc $?"; tail -n 2 gpurun_out/smoke.log | cut -c1-300
timeout 900 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "reference arm rc $?"; cut -c1-200 gpurun_out/bench_ref.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "default bench rc $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_default.log").read().strip().splitlines()[-1])
print("value",round(d["value"],1),"ms",round(d["ms_per_step"],1),"e2e",round(d["e2e"]["value"],1),"cpu",round(d["cpu_baseline"]["value"],3),"launches",d["gpu_launches"], d["clocks"])
PY
