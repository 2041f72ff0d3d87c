#!/bin/bash
# Runs the GPU tests in groups, each group in its own process (a trapped kernel poisons the CUDA
# context); a failing group is re-run one test id per process.  Logs to gpurun_out/.
mkdir -p gpurun_out
LOG=gpurun_out/first_light.log
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
: > $LOG
run_group() {
  local name="$1"; shift
  echo "=== GROUP $name" | tee -a $LOG
  timeout 900 python -m pytest "$@" -q -s -p no:cacheprovider > gpurun_out/group_$name.log 2>&1
  local rc=$?
  echo "--- group $name exit $rc" | tee -a $LOG
  grep -E "max_abs_err|normalised err|passed|failed|PASSED|FAILED|Error|error|timeout|gww" gpurun_out/group_$name.log | head -n 80 >> $LOG
  if [ $rc -ne 0 ]; then
    python -m pytest "$@" --collect-only -q -p no:cacheprovider 2>/dev/null | grep "::" > gpurun_out/ids_$name.txt
    while read -r tid; do
      echo "=== $tid" >> $LOG
      timeout 150 python -m pytest "$tid" -x -q -s -p no:cacheprovider 2>&1 | grep -vE "^\s*$" | tail -n 30 >> $LOG
      echo "--- exit ${PIPESTATUS[0]}" >> $LOG
    done < gpurun_out/ids_$name.txt
  fi
}
run_group gemm tests/test_kernels_gpu.py -m gpu -k "gemm"
run_group misc tests/test_kernels_gpu.py -m gpu -k "layernorm or head or logmel"
run_group attn tests/test_kernels_gpu.py -m gpu -k "attention"
run_group enc tests/test_encoder_gpu.py -m gpu
grep -E "^(=== GROUP|--- group)" $LOG
