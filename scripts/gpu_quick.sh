#!/bin/bash
# quick loop: parity subset, bench stage times, optional ncu --set full of one kernel ($1 regex, $2 tag)
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -q -m gpu -x -k "param_sets or cfg2 or long_protein or degenerate" > gpurun_out/t_quick.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_quick.log
timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print({k:d[k] for k in ['value','ms_per_step','build_ms']}); print({k:round(v,3) for k,v in d['stage_ms'].items()}); print(d['e2e']['ms_per_step'], d['queries']['ms_per_batch'])
PY
tail -3 gpurun_out/bench_quick.err
if [ -n "$1" ]; then
timeout -k 10 1200 ncu --set full --clock-control none --import-source on -k regex:"$1" -s 1 -c 1 -o gpurun_out/prof_$2 -f python scripts/prof_target.py > gpurun_out/ncu_full_$2.log 2>&1
echo "ncu full rc=$?"; tail -1 gpurun_out/ncu_full_$2.log
fi
