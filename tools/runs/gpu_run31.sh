#!/bin/bash
# run 31: SM clock / power while the attention kernel (and each GEMM shape) runs back to back for seconds
mkdir -p gpurun_out
sample() { # $1 = label, rest = command
  local label=$1; shift
  nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.sw_power_cap,temperature.gpu --format=csv,noheader,nounits -lms 200 > gpurun_out/clk_$label.csv 2>/dev/null &
  local pid=$!
  "$@"
  kill $pid 2>/dev/null
  python - "$label" <<'PY'
import sys,statistics
rows=[l.strip().split(", ") for l in open(f"gpurun_out/clk_{sys.argv[1]}.csv") if l.strip()]
rows=[r for r in rows if len(r)>=2]
clk=[float(r[0]) for r in rows]; pw=[float(r[1]) for r in rows]
load=[(c,p) for c,p in zip(clk,pw) if p>400]
if load:
    print(sys.argv[1], "samples under load", len(load), "sm MHz median", statistics.median(c for c,_ in load), "power W median", statistics.median(p for _,p in load))
else:
    print(sys.argv[1], "no loaded samples", len(rows), clk[:5], pw[:5])
PY
}
sample attn timeout 200 python tools/attn_bench.py --reps 2500
sample gemm timeout 200 python tools/gemm_bench.py --reps 1500
