#!/bin/bash
# run 21: warp-elected mbarrier arrivals (attention s_free / p_full, GEMM tmem_empty)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -k "attention or gemm" -q -x -p no:cacheprovider > gpurun_out/k_tests.log 2>&1; rc=$?
echo "kernel tests rc $rc"; tail -n 2 gpurun_out/k_tests.log
if [ $rc -ne 0 ]; then grep -E "gww:|Error" gpurun_out/k_tests.log | head; exit 1; fi
timeout 120 python tools/attn_bench.py
GWW_GEMM_MC=1 timeout 300 python tools/gemm_bench.py
GWW_GEMM_MC=2 timeout 300 python tools/gemm_bench.py
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc $?"
python - <<'PY'
import json
for f in ["gpurun_out/bench.log"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value",round(d["value"],1), "ms",round(d["ms_per_step"],1), "full", d.get("value_full_final_layer"), d["clocks"])
        print("   ", {k:(round(v["ms_per_step"],2), round(v.get("tflops",0))) for k,v in d["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
