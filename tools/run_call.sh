#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_gputest10.log 2>&1; tail -4 gpurun_out/r2_gputest10.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -3 gpurun_out/r2_bench_final.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; tail -2 gpurun_out/r2_bench_reference.err; cut -c1-400 gpurun_out/r2_bench_reference.json
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['roofline_gemm']['frac'])
print({k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
m=d['mlgwsc']; print(m['value'], m['ms'], {k:round(v['ms'],1) for k,v in m['kernels_rank0'].items()})
g=d['glitch_small']; print(g['value'], g['ms_per_step'])
P
