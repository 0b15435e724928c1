#!/bin/bash
# multi-GPU correctness: the sharded build on every visible GPU against the oracle
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout -k 10 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "multi_gpu or lifecycle or cfg1_full" > gpurun_out/t_mg.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/t_mg.log
