#!/bin/bash
# Runs each GPU test file in its own process (a device-side trap only poisons one file) and keeps the logs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in "$@"; do
  b=$(basename "$f" .py)
  timeout 600 python -m pytest "$f" -q -m gpu -s > "gpurun_out/${b}.log" 2>&1
  echo "== $f exit $?"
  tail -n 40 "gpurun_out/${b}.log"
done
