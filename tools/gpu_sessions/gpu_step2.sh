#!/bin/bash
# step-2 path: GPU tests, segment profile, bench.py default line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_step2_gpu.py tests/test_gmmn_fused_gpu.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_step2_gpu.log 2>&1
echo "== step2 tests exit $?"; tail -n 5 gpurun_out/test_step2_gpu.log
timeout 300 python tools/step2_bench.py --steps 6 --warmup 3 --skip-unfused --out gpurun_out/step2_bench.json > gpurun_out/step2_bench.log 2>&1
echo "== step2 bench exit $?"; tail -n 3 gpurun_out/step2_bench.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "== bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print(d['value'], d['e2e']['value'], d['step2'])"
