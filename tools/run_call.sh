#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_qfront_gpu.py tests/test_mlgwsc_golden.py tests/test_train_geometry.py -m gpu -q -x -s 2>&1 | grep -v Warning | grep "spectrogram\|passed\|failed\|Error" | head
python tools/qscan_bench.py 2>&1 | tail -1
