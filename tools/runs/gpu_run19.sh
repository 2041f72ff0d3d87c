#!/bin/bash
# run 19: attention ablation builds (wrong numerics on purpose) to find the binding resource
mkdir -p gpurun_out
: > gpurun_out/attn_ablate.jsonl
timeout 120 python tools/attn_bench.py >> gpurun_out/attn_ablate.jsonl 2>gpurun_out/attn_ablate.err
for ab in 1 2 4 8 6 14; do
  GWW_LIB=$PWD/gw_whisper_b200/variants/libgww_ab$ab.so timeout 120 python tools/attn_bench.py >> gpurun_out/attn_ablate.jsonl 2>>gpurun_out/attn_ablate.err
done
cat gpurun_out/attn_ablate.jsonl
tail -n 3 gpurun_out/attn_ablate.err
