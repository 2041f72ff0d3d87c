#!/bin/bash
# run 47: launch list of the MLGWSC-1 workload (Q front end + whisper-tiny): which Q-Adapter kernel dominates
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_mlgwsc.csv python bench.py --workload mlgwsc --model tiny --batch 256 --steps 1 --warmup 1 > gpurun_out/ncu_mlgwsc.log 2>&1; echo "rc $?"
python tools/summarize_launches.py gpurun_out/launches_mlgwsc.csv "mlgwsc 256 windows" | head -24
