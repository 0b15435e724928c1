#!/bin/bash
# quick: parity subset that exercises query_hits + headline bench only
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -q -m gpu -x -k "${1:-param_sets or api_mirror or sharded_build_in_one_process or kat or cfg3_semi or degenerate}" > gpurun_out/t_quick.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_quick.log
timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print({k:d[k] for k in ['value','ms_per_step','build_ms']}); print({k:round(v,3) for k,v in d['stage_ms'].items()})
print('e2e', d['e2e']['ms_per_step'], d['e2e']['d2h_bytes_per_step'], 'queries', d['queries'], d['parity'])
r=d['roofline']; print(r['kernel'], round(r['frac'],3), round(r['avg_launch_ms'],4), [(o['kernel'],round(o['frac'],3),round(o['avg_launch_ms'],4)) for o in r['other_kernels']])
PY
