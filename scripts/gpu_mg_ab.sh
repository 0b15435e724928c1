#!/bin/bash
# A/B of the cut planner at N GPUs in ONE call (same box): folded slices vs one contiguous slice per rank
N=${1:-8}; K=${2:-5}
mkdir -p gpurun_out
for tag in fold nofold; do
  if [ $tag = nofold ]; then export DBI_MG_FOLD=0; else unset DBI_MG_FOLD; fi
  timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $K --warmup 3 > gpurun_out/bench_mg${N}_$tag.json 2> gpurun_out/bench_mg${N}_$tag.err; echo "$tag rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_mg${N}_$tag.json') if l.startswith('{')][-1])
print('$tag', {k:d[k] for k in ['value','ms_per_step']}, 'e2e', round(d['e2e']['ms_per_step'],2), 'steps', d.get('step_ms_rank0'))
print('  stages', d['host_stage_ms_rank0_last_step'])
for k, v in d['per_rank_last_step'].items(): print('   ', k, v)
print('  split', d.get('split_mass'))
PY
done
