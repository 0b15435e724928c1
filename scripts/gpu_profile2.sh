#!/bin/bash
# ncu --set full of the heaviest kernels of one cfg2 build (second build of prof_target.py)
mkdir -p gpurun_out
R=${1:-r01b}
timeout -k 10 300 python scripts/prof_target.py > gpurun_out/plain2_$R.log 2>&1 &&
timeout -k 10 1200 ncu --set full --clock-control none --import-source on -k regex:"rs_onesweep_kernel|mod_emit_kernel|mod_count_kernel|digest_emit_kernel|digest_count_kernel|rs_histogram_kernel" -s 25 -c 25 -o gpurun_out/prof_$R -f python scripts/prof_target.py > gpurun_out/ncu_full_$R.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full_$R.log
