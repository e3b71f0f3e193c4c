#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gmmn_fused_gpu.py tests/test_step2_gpu.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_fused_v6.log 2>&1
echo "== fused/step2 tests exit $?"; tail -n 4 gpurun_out/test_fused_v6.log
timeout 200 python tools/ncu_new_kernels.py > gpurun_out/new_kernels_timing.json 2> gpurun_out/new_kernels_timing.err
echo "== timings exit $?"; cat gpurun_out/new_kernels_timing.json
