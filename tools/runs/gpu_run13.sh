#!/bin/bash
# run 13: GEMM micro-benchmark (single-CTA vs CTA-pair, lean producer loop) + ncu full capture of each shape
mkdir -p gpurun_out
GWW_GEMM_MC=2 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -k gemm -q -x -p no:cacheprovider > gpurun_out/gemm_pair.log 2>&1; echo "gemm tests (pair) rc $?"
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -k gemm -q -x -p no:cacheprovider > gpurun_out/gemm_auto.log 2>&1; echo "gemm tests (auto) rc $?"
for mc in 1 2; do
  GWW_GEMM_MC=$mc timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench_mc$mc.json 2> gpurun_out/gemm_bench_mc$mc.err
  cat gpurun_out/gemm_bench_mc$mc.json
done
for mc in 1 2; do
  GWW_GEMM_MC=$mc timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 16 -o gpurun_out/prof_r1d_mc$mc python tools/gemm_bench.py --reps 1 > gpurun_out/ncu_gemm_mc$mc.log 2>&1
  tail -n 2 gpurun_out/ncu_gemm_mc$mc.log
done
