#!/bin/bash
# run 27: ncu full capture of the four GEMM shapes with the reworked kernel (auto pair mode)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 4 -o gpurun_out/prof_r1f_gemm python tools/gemm_bench.py --reps 1 --warmup 0 > gpurun_out/ncu_gemm.log 2>&1
tail -n 2 gpurun_out/ncu_gemm.log
