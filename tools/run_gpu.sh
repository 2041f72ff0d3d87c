#!/bin/bash
# helper for gpurun calls: runs the named steps and keeps logs under gpurun_out/ (scratch)
set -x
mkdir -p gpurun_out
"$@"
