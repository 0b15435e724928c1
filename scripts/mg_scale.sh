#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "multi_gpu" > gpurun_out/t_mg8.log 2>&1; echo "pytest(8 gpus) rc=$?"; tail -3 gpurun_out/t_mg8.log
for N in 4 8; do bash scripts/mg_bench.sh $N 2>&1 | grep -v "^\*\|OMP_NUM" | cut -c1-1500; done
