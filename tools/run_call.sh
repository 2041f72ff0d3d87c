#!/bin/bash
# one gpurun call of round 2 (edited per call): logs go to gpurun_out/
mkdir -p gpurun_out
rm -f gpurun_out/r2_attn_tune.jsonl
for v in default $(ls gw_whisper_b200/variants/ | sed 's/libgww_//; s/.so//'); do
  if [ $v = default ]; then unset GWW_LIB; else export GWW_LIB=$PWD/gw_whisper_b200/variants/libgww_$v.so; fi
  python tools/attn_bench.py --reps 10 >> gpurun_out/r2_attn_tune.jsonl 2>> gpurun_out/r2_attn_tune.err
done
unset GWW_LIB
cat gpurun_out/r2_attn_tune.jsonl
timeout 600 python -m pytest tests/test_train_geometry.py tests/test_kernels_gpu.py tests/test_encoder_gpu.py -m gpu -q -s -x > gpurun_out/r2_gputest5.log 2>&1
tail -12 gpurun_out/r2_gputest5.log
python bench.py --workload mlgwsc --mlgwsc-scale 0.1 --no-cpu-baseline > gpurun_out/r2_bench_mlgwsc_small.json 2> gpurun_out/r2_bench_mlgwsc_small.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_mlgwsc_raw.csv python bench.py --workload mlgwsc --mlgwsc-scale 0.03 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
tail -3 gpurun_out/r2_ncu_bench.log | cut -c1-300
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
cut -c1-200 gpurun_out/r2_bench4.json
