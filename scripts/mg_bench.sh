#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_mg$N.json 2> gpurun_out/bench_mg$N.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_mg$N.err | cut -c1-400
python - <<PY
import json
for l in open('gpurun_out/bench_mg$N.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ['value','n_gpus','ms_per_step','queries','all_to_all','roofline','e2e','gpu_launches','clocks','host_stage_ms_rank0_last_step']}); print(d['config'])
PY
