#!/bin/bash
# run 14: ncu full capture of the four GEMM shapes, single-CTA and CTA-pair (one launch each)
mkdir -p gpurun_out
for mc in 1 2; do
  GWW_GEMM_MC=$mc timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 4 -o gpurun_out/prof_r1d_mc$mc python tools/gemm_bench.py --reps 1 --warmup 0 > gpurun_out/ncu_gemm_mc$mc.log 2>&1
  tail -n 2 gpurun_out/ncu_gemm_mc$mc.log
done
ls -la gpurun_out
