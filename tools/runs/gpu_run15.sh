#!/bin/bash
# run 15: GEMM epilogue rework (bias in smem, batched LDS->STG, no divisions): tests, micro-bench, step bench
mkdir -p gpurun_out
GWW_GEMM_MC=2 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -k gemm -q -x -p no:cacheprovider > gpurun_out/gemm_pair.log 2>&1; rc1=$?
GWW_GEMM_MC=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -k gemm -q -x -p no:cacheprovider > gpurun_out/gemm_single.log 2>&1; rc2=$?
echo "gemm tests pair rc $rc1 single rc $rc2"
tail -n 3 gpurun_out/gemm_pair.log gpurun_out/gemm_single.log
if [ $rc1 -ne 0 ] || [ $rc2 -ne 0 ]; then grep -E "max_abs_err|Error|error|gww:" gpurun_out/gemm_pair.log gpurun_out/gemm_single.log | head -20; exit 1; fi
for mc in 1 2; do
  GWW_GEMM_MC=$mc timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench_mc$mc.json 2> gpurun_out/gemm_bench_mc$mc.err
  cat gpurun_out/gemm_bench_mc$mc.json
done
timeout 600 python -m pytest tests/test_encoder_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/enc_tests.log 2>&1; echo "encoder tests rc $?"
tail -n 2 gpurun_out/enc_tests.log
GWW_GEMM_MC=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mc1.log 2> gpurun_out/bench_mc1.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc $?"
tail -n 3 gpurun_out/bench.err
python - <<'PY'
import json
for f in ["gpurun_out/bench_mc1.log","gpurun_out/bench.log"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value",round(d["value"],1), "ms",round(d["ms_per_step"],1), "full", d.get("value_full_final_layer"), d["clocks"])
        print("   ", {k:(round(v["ms_per_step"],2), round(v.get("tflops",0))) for k,v in d["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
