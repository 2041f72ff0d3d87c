#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k attention 2>&1 | tail -5
out=gpurun_out/r2_attn_maxwarps.jsonl; : > $out
python tools/attn_bench.py --reps 20 >> $out 2>gpurun_out/r2_attn_mw.err
for v in mx0 mx1g0 mx1p3 mx1p1; do
  GWW_LIB=gw_whisper_b200/variants/lib_$v.so timeout 120 python tools/attn_bench.py --reps 20 >> $out 2>>gpurun_out/r2_attn_mw.err
done
python tools/attn_bench.py --reps 20 --d 384 >> $out 2>>gpurun_out/r2_attn_mw.err
cat $out
for v in mx1tr; do
GWW_LIB=gw_whisper_b200/variants/lib_$v.so timeout 120 python tools/attn_bench.py --reps 1 --warmup 0 --det-windows 74 2>&1 | grep "gww-" | cut -c1-360 | sort > gpurun_out/r2_attn_trace_$v.txt
cat gpurun_out/r2_attn_trace_$v.txt
done
