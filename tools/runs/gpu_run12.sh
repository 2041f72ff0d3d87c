#!/bin/bash
# run 12: genuine CTA-pair GEMM (tcgen05.mma.cta_group::2, MC=2): correctness first, then A/B in bench
mkdir -p gpurun_out
GWW_GEMM_MC=2 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -k gemm -q -x -s -p no:cacheprovider > gpurun_out/gemm_pair.log 2>&1; rc=$?
echo "gemm tests (pair forced) rc $rc"
grep -E "max_abs_err|passed|failed|Error|error|gww:" gpurun_out/gemm_pair.log | tail -25
if [ $rc -ne 0 ]; then exit 1; fi
GWW_GEMM_MC=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mc1.log 2> gpurun_out/bench_mc1.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc $?"
tail -n 3 gpurun_out/bench.err
python - <<'PY'
import json
for f in ["gpurun_out/bench_mc1.log","gpurun_out/bench.log"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value",round(d["value"],1), "ms",round(d["ms_per_step"],1), "full", d.get("value_full_final_layer"))
        print("   ", {k:(round(v["ms_per_step"],2), round(v.get("tflops",0))) for k,v in d["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
timeout 600 python -m pytest tests/test_encoder_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/enc_tests.log 2>&1; echo "encoder tests rc $?"
tail -n 3 gpurun_out/enc_tests.log
