#!/bin/bash
# run 16: warp-uniform elected issue loops (GEMM + attention), epilogue rework: full GPU tests, micro-bench, step bench
mkdir -p gpurun_out
bash tools/gpu_first_light.sh > gpurun_out/fl_stdout.log 2>&1
grep -E "GROUP|group .* exit" gpurun_out/first_light.log
grep -E "FAILED|Error|timeout" gpurun_out/first_light.log | cut -c1-200 | head
for mc in 1 2; do
  GWW_GEMM_MC=$mc timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench_mc$mc.json 2> gpurun_out/gemm_bench_mc$mc.err
  cat gpurun_out/gemm_bench_mc$mc.json
done
GWW_GEMM_MC=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mc1.log 2> gpurun_out/bench_mc1.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc $?"
tail -n 3 gpurun_out/bench.err
python - <<'PY'
import json
for f in ["gpurun_out/bench_mc1.log","gpurun_out/bench.log"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value",round(d["value"],1), "ms",round(d["ms_per_step"],1), "full", d.get("value_full_final_layer"), d["clocks"])
        print("   ", {k:(round(v["ms_per_step"],2), round(v.get("tflops",0))) for k,v in d["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
