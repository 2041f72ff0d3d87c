#!/bin/bash
# run 57: det-windows per encoder pass (chunk) sweep at the headline workload
mkdir -p gpurun_out
for c in 256 296 512 1024 2048; do
  timeout 300 python bench.py --chunk $c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c$c.log 2> gpurun_out/bench_c$c.err
  python - $c <<'PY'
import json,sys
c=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/bench_c{c}.log").read().strip().splitlines()[-1])
    print("chunk",c,"value",round(d["value"],1),"ms",round(d["ms_per_step"],1),"e2e",round(d["e2e"]["value"],1),d["clocks"]["sm_mhz"], {k:round(v["ms_per_step"],1) for k,v in d["kernels"].items()})
except Exception as e:
    print("chunk",c,"failed",e, open(f"gpurun_out/bench_c{c}.err").read()[-300:])
PY
done
