#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k attention 2>&1 | tail -3
out=gpurun_out/r2_attn_defer.jsonl; : > $out
python tools/attn_bench.py --reps 20 >> $out 2>gpurun_out/r2_attn_mw.err
GWW_LIB=gw_whisper_b200/variants/lib_de0.so timeout 120 python tools/attn_bench.py --reps 20 >> $out 2>>gpurun_out/r2_attn_mw.err
python tools/attn_bench.py --reps 20 >> $out 2>gpurun_out/r2_attn_mw.err
GWW_LIB=gw_whisper_b200/variants/lib_de0.so timeout 120 python tools/attn_bench.py --reps 20 >> $out 2>>gpurun_out/r2_attn_mw.err
python tools/attn_bench.py --reps 20 --d 384 >> $out 2>>gpurun_out/r2_attn_mw.err
cat $out
GWW_LIB=gw_whisper_b200/variants/lib_de1tr.so timeout 120 python tools/attn_bench.py --reps 1 --warmup 0 --det-windows 74 2>&1 | grep "gww-" | cut -c1-360 | sort | head -8
