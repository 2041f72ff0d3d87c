#!/bin/bash
# run 43 (profile evidence for round 1, final kernels): bench first (no profiler), then the ncu launch list of
# the same command, then one ncu --set full capture of the top kernels
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_p.log 2> gpurun_out/bench_p.err; echo "bench rc $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc $?"
python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|attention_persist|layernorm_kernel|logmel_kernel" -s 4 -c 9 -o gpurun_out/prof_r1g python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/ncu_full.log
ls -la gpurun_out | head -20
