#!/bin/bash
# run 11: coalesced GEMM epilogue; 2-CTA weight multicast (forced on for the GEMM tests, A/B in bench)
mkdir -p gpurun_out
GWW_GEMM_MC=2 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -k gemm -q -x -s -p no:cacheprovider > gpurun_out/gemm_mc2.log 2>&1; echo "gemm tests (MC=2 forced) rc $?"
bash tools/gpu_first_light.sh > gpurun_out/fl_stdout.log 2>&1
GWW_GEMM_MC=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mc1.log 2> gpurun_out/bench_mc1.err
GWW_GEMM_MC=2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mc2.log 2> gpurun_out/bench_mc2.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc $?" >> gpurun_out/bench.err
grep -E "passed|failed|Error|error" gpurun_out/gemm_mc2.log | tail -5
grep -E "GROUP|group .* exit" gpurun_out/first_light.log
grep -E "FAILED|Error|timeout" gpurun_out/first_light.log | cut -c1-200 | head
tail -n 3 gpurun_out/bench.err gpurun_out/bench_mc2.err
python - <<'PY'
import json
for f in ["gpurun_out/bench_mc1.log","gpurun_out/bench_mc2.log","gpurun_out/bench.log"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value",round(d["value"],1), "ms",round(d["ms_per_step"],1), "full", d.get("value_full_final_layer"))
        print("   ", {k:(round(v["ms_per_step"],2), round(v.get("tflops",0))) for k,v in d["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
