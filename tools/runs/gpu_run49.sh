#!/bin/bash
# run 49: packed-fp32 GELU epilogue (FFMA2/FMUL2), packed softmax scale/sum: tests + micro-bench + step bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -k "gemm" -q -x -p no:cacheprovider > gpurun_out/k_tests.log 2>&1; rc=$?
echo "gemm tests rc $rc"; tail -n 1 gpurun_out/k_tests.log
if [ $rc -ne 0 ]; then grep -E "max_abs_err|gww:|Error" gpurun_out/k_tests.log | head; exit 1; fi
GWW_GEMM_MC=1 timeout 200 python tools/gemm_bench.py --only fc1 | cut -c1-300
GWW_GEMM_MC=2 timeout 200 python tools/gemm_bench.py --only fc1 | cut -c1-300
timeout 600 python -m pytest tests/test_encoder_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/enc_tests.log 2>&1; echo "encoder+fullsize tests rc $?"; tail -n 1 gpurun_out/enc_tests.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc $?"
python - <<'PY'
import json
for f in ["gpurun_out/bench.log"]:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value",round(d["value"],1), "ms",round(d["ms_per_step"],1), "full", d.get("value_full_final_layer"), d["clocks"])
    print("   ", {k:(round(v["ms_per_step"],2), round(v.get("tflops",0))) for k,v in d["kernels"].items()})
PY
