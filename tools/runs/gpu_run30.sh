#!/bin/bash
# run 30: persistent attention kernel: tests, micro-bench vs per-item kernels, encoder tests, step bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -k "attention" -q -x -s -p no:cacheprovider > gpurun_out/attn_persist.log 2>&1; rc=$?
echo "attention tests (persistent) rc $rc"; grep -E "max_abs_err|passed|failed" gpurun_out/attn_persist.log | tail -7
if [ $rc -ne 0 ]; then grep -E "gww:" gpurun_out/attn_persist.log | sort | uniq -c | head -8; tail -n 5 gpurun_out/attn_persist.log; exit 1; fi
timeout 120 python tools/attn_bench.py
GWW_ATTN_PERSIST=0 timeout 120 python tools/attn_bench.py
timeout 120 python tools/attn_bench.py --det-windows 64 --T 6000
timeout 120 python tools/attn_bench.py --det-windows 1024 --T 384
timeout 600 python -m pytest tests/test_encoder_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/enc_tests.log 2>&1; echo "encoder tests rc $?"
tail -n 2 gpurun_out/enc_tests.log
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
