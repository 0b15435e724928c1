#!/bin/bash
# sharded bench line at N GPUs (torchrun), steps $2
N=${1:-2}; K=${2:-5}
mkdir -p gpurun_out
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $K --warmup 3 > gpurun_out/bench_mg$N.json 2> gpurun_out/bench_mg$N.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_mg$N.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_mg$N.json') if l.startswith('{')][-1])
print({k:d[k] for k in ['value','ms_per_step','n_gpus']}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step']); print('a2a', d['all_to_all']); print('parity', d['parity']); print('steps', d.get('step_ms_rank0'), 'split', d.get('split_mass')); print('queries', d['queries'])
print('stages', d['host_stage_ms_rank0_last_step']); print('e2e stages', d.get('e2e_host_ms_rank0_last_step')); [print('  e2e rank', i, x) for i, x in enumerate(d.get('e2e_per_rank', []))]; [print('  ', k, v) for k, v in d['per_rank_last_step'].items()]
r=d['roofline']; print(r['kernel'], round(r['frac'],3), round(r['avg_launch_ms'],4))
PY
