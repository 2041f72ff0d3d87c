#!/bin/bash
# run 46: out_proj with a 128-wide N tile in the step; fc2 / qkv tile experiments
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/enc_tests.log 2>&1; echo "encoder tests rc $?"
for mc in 1 2; do for bn in 128 256; do GWW_GEMM_MC=$mc timeout 200 python tools/gemm_bench.py --bn $bn --only fc2,qkv | cut -c1-400; done; done
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc $?"
python - <<'PY'
import json
for f in ["gpurun_out/bench.log"]:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value",round(d["value"],1), "ms",round(d["ms_per_step"],1), "full", d.get("value_full_final_layer"), d["clocks"])
    print("   ", {k:(round(v["ms_per_step"],2), round(v.get("tflops",0))) for k,v in d["kernels"].items()})
PY
