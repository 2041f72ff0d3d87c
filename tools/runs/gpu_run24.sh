#!/bin/bash
# run 24: attention with one query tile per CTA (two CTAs per SM) vs two tiles per CTA
mkdir -p gpurun_out
for nt in 1 2; do
  GWW_ATTN_NT=$nt timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -k "attention" -q -x -p no:cacheprovider > gpurun_out/attn_nt$nt.log 2>&1; rc=$?
  echo "attention tests NT=$nt rc $rc"; tail -n 1 gpurun_out/attn_nt$nt.log
  if [ $rc -ne 0 ]; then grep -E "gww:|Error" gpurun_out/attn_nt$nt.log | head -5; fi
  GWW_ATTN_NT=$nt timeout 120 python tools/attn_bench.py
done
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
