#!/bin/bash
mkdir -p gpurun_out
N=8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -3 gpurun_out/r2_bench_${N}gpu.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_8gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
m=d['mlgwsc']; print(m['n_gpus'], m['value'], m['ms'], m['triggers'], m['triggers_per_rank'], m['prefix_check'])
P
