#!/bin/bash
# one gpurun call (edited per call): ncu captures of the final round-2 kernels; raw CSV pages come back, reports stay on the box
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
python tools/profile_mlgwsc.py > gpurun_out/r2_profile_mlgwsc.log 2>&1 || exit 1
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  $NCU -k regex:"$rx" -s $skip -c $cnt -o /tmp/$name "$@" > gpurun_out/$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
}
cap r2_ncu_qfront "qadapter_conv|qscan_tiles|qadapter_pool|qscan_interp" 6 6 python tools/profile_mlgwsc.py
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_svn_raw.csv python tools/profile_step.py --windows 148 --reps 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_mlgwsc_raw.csv python tools/profile_mlgwsc.py > /dev/null 2>&1
ls -la gpurun_out/*raw.csv
