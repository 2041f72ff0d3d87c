#!/bin/bash
# run 9: TMEM ubench, attention stagger x FMA-offload variants, last-token pruning tests
mkdir -p gpurun_out
./tools/ubench/tmem_ld.bin > gpurun_out/ubench_tmem.txt 2>&1
bash tools/gpu_first_light.sh > gpurun_out/fl_stdout.log 2>&1
for v in f0s0 f0s1 f6s1 f3s1; do
  GWW_LIB=$PWD/gw_whisper_b200/variants/libgww_$v.so timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.log 2> gpurun_out/bench_$v.err
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc $?" >> gpurun_out/bench.err
cat gpurun_out/ubench_tmem.txt
grep -E "GROUP|group .* exit" gpurun_out/first_light.log
grep -E "pooled|FAILED|Error" gpurun_out/first_light.log | cut -c1-200
tail -n 3 gpurun_out/bench.err
python - <<'PY'
import json
for f in ["gpurun_out/bench_%s.log"%k for k in ("f0s0","f0s1","f6s1","f3s1")]+["gpurun_out/bench.log"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value",round(d["value"],1), "ms",round(d["ms_per_step"],1), "full", d.get("value_full_final_layer"))
        print("   ", {k:(round(v["ms_per_step"],2), round(v.get("tflops",0))) for k,v in d["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
