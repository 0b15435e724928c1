#!/bin/bash
# ncu --set full of ONE kernel (regex $1, skip $2 launches, capture $3) of prof_target.py -> gpurun_out/prof_$4.ncu-rep
mkdir -p gpurun_out
K=$1; S=${2:-1}; C=${3:-1}; R=${4:-k}
timeout -k 10 300 python scripts/prof_target.py > gpurun_out/plain_$R.log 2>&1 &&
timeout -k 10 1200 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c $C -o gpurun_out/prof_$R -f python scripts/prof_target.py > gpurun_out/ncu_full_$R.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full_$R.log
