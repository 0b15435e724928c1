#!/bin/bash
# multi-process sharded build (torchrun, NCCL small collectives, CUDA IPC windows) against the oracle
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout -k 10 ${1:-400} python -m pytest tests -q -m gpu -x -k "multi_gpu_sharded" > gpurun_out/t_mg.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/t_mg.log
