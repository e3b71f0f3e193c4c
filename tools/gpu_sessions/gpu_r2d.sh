#!/bin/bash
# round 2, pass d (2 GPUs): the whole GPU suite (as the driver runs it) with per-test durations
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu --tb=short --durations=15 -p no:cacheprovider > gpurun_out/r02d_gpu_suite.log 2>&1
echo "== gpu suite exit $?"; tail -n 45 gpurun_out/r02d_gpu_suite.log
