#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_qfront_gpu.py tests/test_mlgwsc_golden.py -m gpu -q -x -s 2>&1 | grep -v Warning | grep "qscan\|QScan\|tiles\|spectrogram\|passed\|failed\|Error" | head
python tools/qscan_bench.py 2>&1 | tail -1
python bench.py --workload mlgwsc --no-cpu-baseline > gpurun_out/r2_bench_mlgwsc5.json 2> gpurun_out/r2_bench_mlgwsc5.err; tail -3 gpurun_out/r2_bench_mlgwsc5.err; python - <<'P'
import json
m=json.loads(open('gpurun_out/r2_bench_mlgwsc5.json').read().strip().splitlines()[-1])
print(m['value'], m['ms'], m['triggers'], {k:round(v['ms'],1) for k,v in m['kernels_rank0'].items()})
P
