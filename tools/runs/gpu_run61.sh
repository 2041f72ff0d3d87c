#!/bin/bash
# run 61: 16 (12 for BN=192) epilogue warps for the GELU epilogue vs 8
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -k "gemm" -q -x -p no:cacheprovider > gpurun_out/k_tests.log 2>&1; rc=$?
echo "gemm tests rc $rc"; tail -n 1 gpurun_out/k_tests.log
if [ $rc -ne 0 ]; then grep -E "max_abs_err|gww:|Error" gpurun_out/k_tests.log | head; exit 1; fi
timeout 200 python tools/gemm_bench.py --only fc1 | cut -c1-300
GWW_LIB=$PWD/gw_whisper_b200/variants/libgww_gelu8.so timeout 200 python tools/gemm_bench.py --only fc1 | cut -c1-300
timeout 600 python -m pytest tests/test_encoder_gpu.py -m gpu -q -x -k "not alternative" -p no:cacheprovider > gpurun_out/enc_tests.log 2>&1; echo "encoder tests rc $?"; tail -n 1 gpurun_out/enc_tests.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc $?"
GWW_LIB=$PWD/gw_whisper_b200/variants/libgww_gelu8.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gelu8.log 2> gpurun_out/bench_gelu8.err
python - <<'PY'
import json
for f in ["gpurun_out/bench.log","gpurun_out/bench_gelu8.log"]:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value",round(d["value"],1), "ms",round(d["ms_per_step"],1), d["clocks"]["sm_mhz"], {k:round(v["ms_per_step"],1) for k,v in d["kernels"].items() if k.startswith("gemm")})
PY
