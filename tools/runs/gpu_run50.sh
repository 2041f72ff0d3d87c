#!/bin/bash
# run 50: LayerNorm folded into the neighbouring GEMMs: all GPU tests, then the step bench with and without the fold
mkdir -p gpurun_out
bash tools/gpu_first_light.sh > gpurun_out/fl_stdout.log 2>&1
grep -E "GROUP|group .* exit" gpurun_out/first_light.log
grep -E "FAILED|Error|timeout|gww:" gpurun_out/first_light.log | cut -c1-220 | head
grep -E "max_abs_err" gpurun_out/group_enc.log | cut -c1-200
timeout 900 python -m pytest tests/test_fullsize_gpu.py tests/test_qfront_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/more_tests.log 2>&1; echo "fullsize+qfront rc $?"; tail -n 1 gpurun_out/more_tests.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc $?"
GWW_LN_FOLD=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nofold.log 2> gpurun_out/bench_nofold.err
python - <<'PY'
import json
for f in ["gpurun_out/bench.log","gpurun_out/bench_nofold.log"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value",round(d["value"],1), "ms",round(d["ms_per_step"],1), "full", d.get("value_full_final_layer"), d["clocks"])
        print("   ", {k:(round(v["ms_per_step"],2), round(v.get("tflops",0))) for k,v in d["kernels"].items()})
    except Exception as e:
        print(f, "failed", e); print(open(f.replace(".log",".err")).read()[-500:])
PY
