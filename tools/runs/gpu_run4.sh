#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_first_light.sh > gpurun_out/fl_stdout.log 2>&1
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc $?" >> gpurun_out/bench.err
timeout 300 python bench.py --steps 2 --warmup 3 --chunk 128 --no-cpu-baseline > gpurun_out/bench_chunk128.log 2> gpurun_out/bench_chunk128.err
grep -E "GROUP|group .* exit" gpurun_out/first_light.log
tail -n 2 gpurun_out/smoke.log
tail -n 3 gpurun_out/bench.err
python - <<'PY'
import json
for f in ("gpurun_out/bench.log","gpurun_out/bench_chunk128.log"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value",round(d["value"],1), "ms",round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],1), "host_issue", round(d["host_issue_ms_per_step"],1), "gemm_frac", round(d["roofline"]["frac"],3))
        print("   ", {k:(round(v["ms_per_step"],2), round(v.get("tflops",0))) for k,v in d["kernels"].items()})
        print("    cpu", d.get("cpu_baseline"), d.get("clocks"))
    except Exception as e:
        print(f, "failed", e)
PY
