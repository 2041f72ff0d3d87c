#!/bin/bash
# run 18: ncu full capture of attention v4 (one launch, whisper-base chunk of 256 det-windows)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 2 -c 1 -o gpurun_out/prof_r1e_attn python tools/profile_step.py > gpurun_out/ncu_attn.log 2>&1
tail -n 3 gpurun_out/ncu_attn.log
ls -la gpurun_out
