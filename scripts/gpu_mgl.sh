#!/bin/bash
# sharded build inside one process (world handles on cuda:0) + the rest of the parity suite
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -q -m gpu -x -k "sharded_build_in_one_process" > gpurun_out/t_mgl.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/t_mgl.log
