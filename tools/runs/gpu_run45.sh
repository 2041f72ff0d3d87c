#!/bin/bash
# run 45: full-size property tests; out_proj / fc1 with a 128-wide N tile
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/fullsize.log 2>&1; echo "fullsize tests rc $?"; tail -n 12 gpurun_out/fullsize.log | cut -c1-300
for mc in 1 2; do for bn in 128 256; do GWW_GEMM_MC=$mc timeout 200 python tools/gemm_bench.py --bn $bn --only out_proj,fc1 | cut -c1-400; done; done
