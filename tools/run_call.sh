#!/bin/bash
# one gpurun call of round 2 (edited per call): logs go to gpurun_out/
mkdir -p gpurun_out
rm -f gpurun_out/r2_attn_tune2.jsonl
for v in default $(ls gw_whisper_b200/variants/ | sed 's/libgww_//; s/.so//'); do
  if [ $v = default ]; then unset GWW_LIB; else export GWW_LIB=$PWD/gw_whisper_b200/variants/libgww_$v.so; fi
  python tools/attn_bench.py --reps 10 >> gpurun_out/r2_attn_tune2.jsonl 2>> gpurun_out/r2_attn_tune2.err
done
unset GWW_LIB
cat gpurun_out/r2_attn_tune2.jsonl
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r2_gputest6.log 2>&1
tail -12 gpurun_out/r2_gputest6.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
cut -c1-200 gpurun_out/r2_bench5.json
