#!/bin/bash
# full GPU parity suite + the default bench line (headline + cfg1 / cfg3 / cfg5 sub-results + CPU baseline)
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t_gpu.log
timeout -k 10 900 python bench.py --steps ${1:-10} --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_full.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_full.json'))
print({k:d[k] for k in ['value','ms_per_step','build_ms']}); print({k:round(v,3) for k,v in d['stage_ms'].items()})
print('e2e', d['e2e']); print('queries', d['queries']); print('parity', d['parity']); print('cpu', d['cpu_baseline'])
r=d['roofline']; print(r['kernel'], round(r['frac'],3), round(r['avg_launch_ms'],4), [(o['kernel'],round(o['frac'],3),round(o['avg_launch_ms'],4)) for o in r['other_kernels']])
for k,v in d['configs'].items():
    if isinstance(v,dict):
        print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a not in ('stage_ms','roofline','workload','what')})
        if 'roofline' in v and v['roofline']: print('   roof', v['roofline']['kernel'], round(v['roofline']['frac'],3))
    else: print(k, v)
PY
