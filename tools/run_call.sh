#!/bin/bash
# one gpurun call (edited per call).  This version: the ncu captures behind profiles/r2_ncu_summary.csv,
# r2_ncu_traffic.json and r2_launches_*.csv (raw CSV pages come back in gpurun_out/, reports stay on the box);
# summarise with  python tools/make_ncu_summary.py r2  and  tools/summarize_launches.py.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
python tools/profile_step.py --windows 128 > gpurun_out/r2_profile_step.log 2>&1 || exit 1
python tools/profile_mlgwsc.py > gpurun_out/r2_profile_mlgwsc.log 2>&1 || exit 1
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  $NCU -k regex:"$rx" -s $skip -c $cnt -o /tmp/$name "$@" > gpurun_out/$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
}
cap r2_ncu_attn attention_persist_kernel 1 1 python tools/profile_step.py --windows 128
cp /tmp/r2_ncu_attn.ncu-rep gpurun_out/
cap r2_ncu_logmel logmel_kernel 1 1 python tools/profile_step.py --windows 128
cap r2_ncu_gemm gemm_tc_kernel 6 4 python tools/profile_step.py --windows 128
cap r2_ncu_qfront "qadapter_conv|qscan_tiles|qadapter_pool|qscan_interp" 6 6 python tools/profile_mlgwsc.py
# launch lists (per-launch durations) of one step of each workload
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_svn_raw.csv python tools/profile_step.py --windows 148 --reps 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_mlgwsc_raw.csv python tools/profile_mlgwsc.py > /dev/null 2>&1
ls -la gpurun_out/*raw.csv gpurun_out/*.ncu-rep
