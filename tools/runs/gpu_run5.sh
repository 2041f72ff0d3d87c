#!/bin/bash
# round-1 run 5: 8-warp GEMM epilogue + L2 residual prefetch; pipe micro-benchmark; ncu full capture
mkdir -p gpurun_out
./tools/ubench/pipes.bin > gpurun_out/ubench_pipes.txt 2>&1
bash tools/gpu_first_light.sh > gpurun_out/fl_stdout.log 2>&1
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc $?" >> gpurun_out/bench.err
python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
python tools/profile_step.py > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernelILi256ELi[12]|attention_tc" -s 8 -c 4 -o gpurun_out/prof_r1b python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1
grep -E "GROUP|group .* exit" gpurun_out/first_light.log
tail -n 2 gpurun_out/smoke.log
tail -n 3 gpurun_out/bench.err
python - <<'PY'
import json
for f in ("gpurun_out/bench.log",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value",round(d["value"],1), "ms",round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],1), "host_issue", round(d["host_issue_ms_per_step"],1), "gemm_frac", round(d["roofline"]["frac"],3))
        print("   ", {k:(round(v["ms_per_step"],2), round(v.get("tflops",0))) for k,v in d["kernels"].items()})
        print("    cpu", d.get("cpu_baseline"), d.get("clocks"))
    except Exception as e:
        print(f, "failed", e)
PY
grep -E "threads=512" gpurun_out/ubench_pipes.txt
tail -n 3 gpurun_out/ncu_full.log
