#!/bin/bash
# run 22: dependency-throttled exponential groups (group size x lookahead sweep), attention alone
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -k "attention" -q -x -p no:cacheprovider > gpurun_out/k_tests.log 2>&1; rc=$?
echo "attention tests rc $rc"; tail -n 2 gpurun_out/k_tests.log
: > gpurun_out/attn_groups.jsonl
timeout 120 python tools/attn_bench.py >> gpurun_out/attn_groups.jsonl
for v in g0l2 g4l2 g4l4 g8l1 g8l3 g16l1 g16l2; do
  GWW_LIB=$PWD/gw_whisper_b200/variants/libgww_$v.so timeout 120 python tools/attn_bench.py >> gpurun_out/attn_groups.jsonl
done
cat gpurun_out/attn_groups.jsonl
