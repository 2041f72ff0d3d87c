#!/bin/bash
# run 8: attention v3 (concurrent exps, FMA-pipe offload sweep), logmel 3-frame DFT, mlgwsc bench
mkdir -p gpurun_out
bash tools/gpu_first_light.sh > gpurun_out/fl_stdout.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc $?" >> gpurun_out/bench.err
for k in 0 4 8 10; do
  GWW_LIB=$PWD/gw_whisper_b200/variants/libgww_fma$k.so timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fma$k.log 2> gpurun_out/bench_fma$k.err
done
timeout 600 python bench.py --workload mlgwsc --steps 2 --warmup 3 > gpurun_out/bench_mlgwsc.log 2> gpurun_out/bench_mlgwsc.err; echo "mlgwsc rc $?" >> gpurun_out/bench_mlgwsc.err
grep -E "GROUP|group .* exit" gpurun_out/first_light.log
grep -E "logmel window|attn" gpurun_out/first_light.log | cut -c1-160
tail -n 3 gpurun_out/bench.err gpurun_out/bench_mlgwsc.err
python - <<'PY'
import json
for f in ["gpurun_out/bench.log"]+["gpurun_out/bench_fma%d.log"%k for k in (0,4,8,10)]+["gpurun_out/bench_mlgwsc.log"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value",round(d["value"],1), "ms",round(d["ms_per_step"],1))
        print("   ", {k:(round(v["ms_per_step"],2), round(v.get("tflops",0))) for k,v in d["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
