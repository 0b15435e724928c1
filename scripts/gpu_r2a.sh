#!/bin/bash
# round-2 iteration: GPU parity suite + bench (no CPU baseline) + stage times
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/t_gpu.log
timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_iter.json'))
print({k:d[k] for k in ['value','ms_per_step','build_ms']}); print({k:round(v,3) for k,v in d['stage_ms'].items()}); print(d['e2e']['ms_per_step'], d['queries']['ms_per_batch'])
r=d['roofline']; print(r['kernel'], r['frac'], r['avg_launch_ms'], [(o['kernel'],o['frac'],o['avg_launch_ms']) for o in r['other_kernels']])
PY
tail -3 gpurun_out/bench_iter.err
