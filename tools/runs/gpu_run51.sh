#!/bin/bash
# run 51: alternative-path agreement test, default bench (with CPU baseline), ncu --set full of the final kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_encoder_gpu.py -m gpu -k alternative -q -x -s -p no:cacheprovider > gpurun_out/alt_tests.log 2>&1; echo "alternative-path tests rc $?"; grep -E "max \|alt|passed|failed" gpurun_out/alt_tests.log | tail -8
timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "default bench rc $?"
python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|attention_persist" -s 4 -c 7 -o gpurun_out/prof_r1h python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/ncu_full.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_default.log").read().strip().splitlines()[-1])
print("value",round(d["value"],1),"ms",round(d["ms_per_step"],1),"e2e",round(d["e2e"]["value"],1),"cpu",round(d["cpu_baseline"]["value"],3),d["cpu_baseline"]["sample"][:60],"launches",d["gpu_launches"], d["clocks"])
print({k:(round(v["ms_per_step"],2), round(v.get("tflops",0))) for k,v in d["kernels"].items()})
print("roofline", {k:d["roofline"][k] for k in ("bound","achieved","peak","frac","traffic")})
print("roofline_gemm", {k:d["roofline_gemm"][k] for k in ("bound","achieved","peak","frac")})
PY
