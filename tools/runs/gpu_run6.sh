#!/bin/bash
# round-1 run 6: first light of front end B (QScan + Q-Adapter + GWWhisperClassifier)
mkdir -p gpurun_out
run_ids() {
  local name="$1"; shift
  python -m pytest "$@" --collect-only -q -p no:cacheprovider 2>/dev/null | grep "::" > gpurun_out/ids_$name.txt
  : > gpurun_out/q_$name.log
  while read -r tid; do
    echo "=== $tid" >> gpurun_out/q_$name.log
    timeout 300 python -m pytest "$tid" -x -q -s -p no:cacheprovider 2>&1 | grep -vE "^\s*$" | tail -n 25 >> gpurun_out/q_$name.log
    echo "--- exit ${PIPESTATUS[0]}" >> gpurun_out/q_$name.log
  done < gpurun_out/ids_$name.txt
}
run_ids qfront tests/test_qfront_gpu.py -m gpu
grep -E "^===|^--- exit|error|Error|normalised|max_abs_err|planes|passed|failed|gww" gpurun_out/q_qfront.log | cut -c1-220
