#!/bin/bash
# quick iteration loop: parity subset + e2e breakdown + bench
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t_gpu.log
DBI_TRACE=1 timeout -k 10 300 python scripts/e2e_breakdown.py > gpurun_out/e2e.log 2>&1; echo "e2e rc=$?"; cat gpurun_out/e2e.log | tail -30
timeout -k 10 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_iter.json'))
print({k:d[k] for k in ['value','ms_per_step','build_ms','stage_ms','roofline','e2e','queries','gpu_launches','clocks']})
PY
tail -3 gpurun_out/bench_iter.err
