#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_first_light.sh > gpurun_out/fl_stdout.log 2>&1
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc $?" >> gpurun_out/bench.err
for c in 32 64 128; do
  timeout 300 python bench.py --steps 2 --warmup 3 --chunk $c --no-cpu-baseline > gpurun_out/bench_chunk$c.log 2> gpurun_out/bench_chunk$c.err
done
python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
python tools/profile_step.py > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernelILi256ELi1|attention_tc" -s 6 -c 2 -o gpurun_out/prof_r1 python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1
grep -E "GROUP|group .* exit" gpurun_out/first_light.log
tail -2 gpurun_out/smoke.log
cat gpurun_out/bench.log
for c in 32 64 128; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_chunk$c.log").read().strip().splitlines()[-1])
    print("chunk $c", d["value"], d["ms_per_step"], d["host_issue_ms_per_step"], {k:round(v["ms_per_step"],2) for k,v in d["kernels"].items()})
except Exception as e:
    print("chunk $c failed", e)
PY
done
tail -3 gpurun_out/ncu_launches.log gpurun_out/ncu_full.log
