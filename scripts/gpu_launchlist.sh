#!/bin/bash
mkdir -p gpurun_out
R=${1:-x}
timeout -k 10 300 python scripts/prof_target.py > gpurun_out/plain_$R.log 2>&1 &&
timeout -k 10 900 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv python scripts/prof_target.py > gpurun_out/ncu_launch_$R.log 2>&1
echo "rc=$?"
