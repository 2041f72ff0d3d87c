#!/bin/bash
# run 17: attention v4 (four softmax warpgroups, key halves): tests then step bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -k attention -q -x -s -p no:cacheprovider > gpurun_out/attn_tests.log 2>&1; rc=$?
echo "attention tests rc $rc"; grep -E "max_abs_err|passed|failed|gww:" gpurun_out/attn_tests.log | tail -12
if [ $rc -ne 0 ]; then tail -n 30 gpurun_out/attn_tests.log; exit 1; fi
timeout 600 python -m pytest tests/test_encoder_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/enc_tests.log 2>&1; echo "encoder tests rc $?"
tail -n 2 gpurun_out/enc_tests.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc $?"
tail -n 3 gpurun_out/bench.err
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
