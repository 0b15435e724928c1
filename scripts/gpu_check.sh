#!/bin/bash
# First-contact GPU check: smoke, then the GPU parity suite; every step under its own timeout so a
# hung kernel cannot hold the box.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout -k 10 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== sort hook"; timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "radix_sort_hook or kat_albumin" > gpurun_out/t_sort.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_sort.log
echo "== full gpu suite"; timeout -k 10 900 python -m pytest tests -q -m gpu --durations=20 > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/t_gpu.log
