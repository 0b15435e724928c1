#!/bin/bash
# bench line + ncu launch list + one --set full capture of the dominant kernel (one GPU).
mkdir -p gpurun_out
R=${1:-r01}
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench rc=$?"; cat gpurun_out/bench_$R.json; tail -3 gpurun_out/bench_$R.err
timeout -k 10 300 python scripts/prof_target.py > gpurun_out/plain_$R.log 2>&1 &&
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv python scripts/prof_target.py > gpurun_out/ncu_launch_$R.log 2>&1
echo "ncu launches rc=$?"; tail -2 gpurun_out/plain_$R.log
timeout -k 10 300 python scripts/prof_target.py > gpurun_out/plain2_$R.log 2>&1 &&
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:rs_onesweep -s 29 -c 3 -o gpurun_out/prof_onesweep_$R -f python scripts/prof_target.py > gpurun_out/ncu_full_$R.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full_$R.log
