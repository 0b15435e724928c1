#!/bin/bash
# round evidence: full GPU parity suite, smoke, bench (N=1) with the CPU baseline, --impl reference,
# the ncu launch list of the same command and one ncu --set full capture of the top kernels
mkdir -p gpurun_out
R=${1:-r02}
timeout -k 10 1500 python -m pytest tests -q -m gpu > gpurun_out/t_gpu_$R.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_gpu_$R.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$R.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$R.log
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench rc=$?"
timeout -k 10 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err; echo "bench ref rc=$?"
timeout -k 10 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_bench_$R.log 2>&1; echo "ncu list rc=$?"
timeout -k 10 1200 ncu --set full --clock-control none --import-source on -k regex:"grp_expand_tab_kernel|rs_onesweep_kernel|grp_emit_kernel|grp_count_kernel|digest_emit_kernel|digest_count_kernel|rs_histogram_kernel|site_masks_kernel|dedup_flags_kernel|dedup_emit_kernel|hash_records_kernel|query_kernel|hits_count_runs_kernel|hits_runs_kernel|peps_gather_kernel" -c 90 -o gpurun_out/prof_$R -f python scripts/prof_target.py 20000 1 > gpurun_out/ncu_full_$R.log 2>&1; echo "ncu full rc=$?"
