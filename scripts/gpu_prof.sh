#!/bin/bash
# ncu: launch list of scripts/prof_target.py + --set full capture (with source) of the kernels matching $1; tag $2
mkdir -p gpurun_out
K=$1; R=${2:-k}; C=${3:-3}
timeout -k 10 300 python scripts/prof_target.py > gpurun_out/plain_$R.log 2>&1 &&
timeout -k 10 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_$R.csv python scripts/prof_target.py > gpurun_out/ncu_list_$R.log 2>&1
echo "ncu list rc=$?"
timeout -k 10 1200 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 1 -c $C -o gpurun_out/prof_$R -f python scripts/prof_target.py > gpurun_out/ncu_full_$R.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full_$R.log
