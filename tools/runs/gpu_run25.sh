#!/bin/bash
# run 25: contract check -- default bench (with CPU baseline), reference arm, smoke
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err ) 2> gpurun_out/bench_default.time; echo "default bench rc $?"
tail -n 3 gpurun_out/bench_default.time
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err ) 2> gpurun_out/bench_ref.time; echo "reference arm rc $?"
tail -n 3 gpurun_out/bench_ref.time
cat gpurun_out/bench_ref.log | cut -c1-900
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc $?"; tail -n 3 gpurun_out/smoke.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_default.log").read().strip().splitlines()[-1])
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"cpu",d["cpu_baseline"],"launches",d["gpu_launches"], d["clocks"])
print({k:(round(v["ms_per_step"],2), round(v.get("tflops",0))) for k,v in d["kernels"].items()})
print("roofline", {k:d["roofline"][k] for k in ("bound","achieved","peak","frac","traffic")})
print("roofline_gemm", {k:d["roofline_gemm"][k] for k in ("bound","achieved","peak","frac")})
PY
