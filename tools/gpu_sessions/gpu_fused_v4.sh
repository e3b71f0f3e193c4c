#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gmmn_fused_gpu.py tests/test_step2_gpu.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_fused_v4.log 2>&1
echo "== fused/step2 tests exit $?"; tail -n 4 gpurun_out/test_fused_v4.log
timeout 200 python tools/ncu_new_kernels.py > gpurun_out/new_kernels_timing.json 2> gpurun_out/new_kernels_timing.err
echo "== timings exit $?"; cat gpurun_out/new_kernels_timing.json
timeout 300 python tools/step2_bench.py --steps 6 --warmup 3 --skip-unfused --out gpurun_out/step2_bench.json > gpurun_out/step2_bench.log 2>&1
echo "== step2 bench exit $?"; tail -n 2 gpurun_out/step2_bench.log | cut -c1-900
