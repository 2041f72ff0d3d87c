#!/bin/bash
# run 42: full GPU test suite + step bench with the persistent attention kernel (exp groups 4x2)
mkdir -p gpurun_out
bash tools/gpu_first_light.sh > gpurun_out/fl_stdout.log 2>&1
grep -E "GROUP|group .* exit" gpurun_out/first_light.log
grep -E "FAILED|Error|timeout" gpurun_out/first_light.log | cut -c1-200 | head
timeout 900 python -m pytest tests/test_qfront_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/q_qfront.log 2>&1; echo "qfront tests rc $?"; tail -n 2 gpurun_out/q_qfront.log
timeout 120 python tools/attn_bench.py
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc $?"
GWW_ATTN_PERSIST=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_np.log 2> gpurun_out/bench_np.err
python - <<'PY'
import json
for f in ["gpurun_out/bench.log", "gpurun_out/bench_np.log"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value",round(d["value"],1), "ms",round(d["ms_per_step"],1), "full", d.get("value_full_final_layer"), d["clocks"])
        print("   ", {k:(round(v["ms_per_step"],2), round(v.get("tflops",0))) for k,v in d["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
