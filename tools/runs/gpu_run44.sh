#!/bin/bash
# run 44: the other configurations: MLGWSC-1 (Q front end + tiny), svn tiny, glitch-sized small
mkdir -p gpurun_out
timeout 300 python bench.py --workload mlgwsc --model tiny --batch 1024 --steps 3 --warmup 3 > gpurun_out/bench_mlgwsc.log 2> gpurun_out/bench_mlgwsc.err; echo "mlgwsc rc $?"
timeout 300 python bench.py --model tiny --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tiny.log 2> gpurun_out/bench_tiny.err; echo "tiny rc $?"
timeout 300 python bench.py --model small --batch 512 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small.log 2> gpurun_out/bench_small.err; echo "small rc $?"
python - <<'PY'
import json
for f in ["gpurun_out/bench_mlgwsc.log","gpurun_out/bench_tiny.log","gpurun_out/bench_small.log"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value",round(d["value"],1), "ms",round(d["ms_per_step"],1), "model_tflops", round(d.get("model_tflops",0)))
        print("   ", {k:(round(v["ms_per_step"],2), round(v.get("tflops",0))) for k,v in d["kernels"].items()})
    except Exception as e:
        print(f, "failed", e); print(open(f.replace(".log",".err")).read()[-600:])
PY
