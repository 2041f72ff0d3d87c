#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_encoder_gpu.py tests/test_parity_configs_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x > gpurun_out/r2_gputest9.log 2>&1; tail -4 gpurun_out/r2_gputest9.log
python bench.py --no-cpu-baseline > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; tail -3 gpurun_out/r2_bench8.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
print({k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
m=d['mlgwsc']; print(m['value'], m['ms'], {k:round(v['ms'],1) for k,v in m['kernels_rank0'].items()})
g=d['glitch_small']; print(g['value'], g['ms_per_step'])
P
