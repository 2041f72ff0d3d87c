#!/bin/bash
# run 55: ncu launch list of the final code (same bench command, first 700 launches ~ two steps)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc $?"
python tools/summarize_launches.py gpurun_out/launches.csv "final" | head -16
