#!/bin/bash
# run 10: ncu full capture of attention v3 + the three big GEMM shapes (lts / tensor utilisation)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_encoder_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/enc_tests.log 2>&1; echo "enc tests rc $?"
python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|attention_tc" -s 6 -c 8 -o gpurun_out/prof_r1c python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/enc_tests.log gpurun_out/ncu_full.log
