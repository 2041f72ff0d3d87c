#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r2_logmel_phases.jsonl; : > $out
python tools/logmel_bench.py >> $out 2>&1
for v in lm1 lm2 lm4 lm6 lm7; do GWW_LIB=gw_whisper_b200/variants/lib_$v.so python tools/logmel_bench.py >> $out 2>&1; done
cat $out
