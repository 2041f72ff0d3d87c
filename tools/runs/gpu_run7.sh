#!/bin/bash
# run 7: attention v2 (single-pass softmax, FMA-pipe exp offload)
mkdir -p gpurun_out
bash tools/gpu_first_light.sh > gpurun_out/fl_stdout.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc $?" >> gpurun_out/bench.err
grep -E "GROUP|group .* exit" gpurun_out/first_light.log
grep -E "attn" gpurun_out/first_light.log | cut -c1-200
tail -n 3 gpurun_out/bench.err
python - <<'PY'
import json
for f in ("gpurun_out/bench.log",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value",round(d["value"],1), "ms",round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],1), "host_issue", round(d["host_issue_ms_per_step"],1), "gemm_frac", round(d["roofline"]["frac"],3))
        print("   ", {k:(round(v["ms_per_step"],2), round(v.get("tflops",0))) for k,v in d["kernels"].items()})
        print("    clocks", d.get("clocks"))
    except Exception as e:
        print(f, "failed", e)
PY
