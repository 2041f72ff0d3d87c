#!/bin/bash
# run 33: persistent attention with S prefetch across tiles
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -k "attention" -q -x -s -p no:cacheprovider > gpurun_out/attn_persist.log 2>&1; rc=$?
echo "attention tests rc $rc"; grep -E "max_abs_err|passed|failed" gpurun_out/attn_persist.log | tail -7
if [ $rc -ne 0 ]; then grep -E "gww:" gpurun_out/attn_persist.log | sort | uniq -c | head -8; tail -n 5 gpurun_out/attn_persist.log; exit 1; fi
timeout 120 python tools/attn_bench.py
GWW_LIB=$PWD/gw_whisper_b200/variants/libgww_trace.so timeout 100 python tools/attn_bench.py --reps 3 --warmup 1 2>&1 | grep gww-trace | tail -3
timeout 120 python tools/attn_bench.py --det-windows 64 --T 6000
