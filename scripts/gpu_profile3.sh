#!/bin/bash
# ncu --set full of the group-path kernels of one cfg2 build (second build of prof_target.py)
mkdir -p gpurun_out
R=${1:-r01c}
timeout -k 10 300 python scripts/prof_target.py > gpurun_out/plain3_$R.log 2>&1 &&
timeout -k 10 1200 ncu --set full --clock-control none --import-source on -k regex:"grp_expand_kernel|grp_count_kernel|grp_emit_kernel|digest_emit_kernel|rs_onesweep_kernel" -s 18 -c 18 -o gpurun_out/prof_$R -f python scripts/prof_target.py > gpurun_out/ncu_full_$R.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full_$R.log
