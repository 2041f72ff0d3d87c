#!/bin/bash
# run 60: pair mode re-check per shape after the epilogue changes (forced single / forced pair vs auto), same box
mkdir -p gpurun_out
for mc in auto 1 2; do
  if [ $mc = auto ]; then unset GWW_GEMM_MC; else export GWW_GEMM_MC=$mc; fi
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mc$mc.log 2> gpurun_out/bench_mc$mc.err
  python - $mc <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/bench_mc{sys.argv[1]}.log").read().strip().splitlines()[-1])
print("mc",sys.argv[1],"ms",round(d["ms_per_step"],1),d["clocks"]["sm_mhz"], {k:round(v["ms_per_step"],1) for k,v in d["kernels"].items() if k.startswith("gemm")})
PY
done
