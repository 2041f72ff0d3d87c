#!/bin/bash
# one gpurun call of round 2 (edited per call): logs go to gpurun_out/
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
python tools/profile_step.py --windows 128 > gpurun_out/r2_profile_step.log 2>&1 || exit 1
python tools/profile_mlgwsc.py > gpurun_out/r2_profile_mlgwsc.log 2>&1 || exit 1
$NCU -k regex:attention_persist_kernel -s 1 -c 1 -o gpurun_out/r2_ncu_attn python tools/profile_step.py --windows 128 > gpurun_out/r2_ncu_attn.log 2>&1
$NCU -k regex:logmel_kernel -s 1 -c 1 -o gpurun_out/r2_ncu_logmel python tools/profile_step.py --windows 128 > gpurun_out/r2_ncu_logmel.log 2>&1
$NCU -k regex:gemm_tc_kernel -s 6 -c 4 -o gpurun_out/r2_ncu_gemm python tools/profile_step.py --windows 128 > gpurun_out/r2_ncu_gemm.log 2>&1
$NCU -k regex:"qadapter_conv|qscan_tiles|qadapter_pool|qscan_interp" -s 6 -c 6 -o gpurun_out/r2_ncu_qfront python tools/profile_mlgwsc.py > gpurun_out/r2_ncu_qfront.log 2>&1
ls -la gpurun_out/*.ncu-rep
