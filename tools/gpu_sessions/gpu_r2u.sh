#!/bin/bash
# round 2, pass u (1 GPU): graph-generator updates in the fused work-list kernel -- tests, config-5 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_step2_gpu.py tests/test_gmmn_fused_gpu.py tests/test_gmmn_gpu.py tests/test_graph_gpu.py -q -m gpu -s --tb=short -p no:cacheprovider > gpurun_out/r02u_tests.log 2>&1
echo "== tests exit $?"; grep -E "gcn generator|passed|failed|Error" gpurun_out/r02u_tests.log | tail -n 14
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-library-baseline --no-parity --no-numerics --no-cpu-baseline --no-transforms > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02u_bench.json'))
print('value', round(d['value'], 1), 'step2', round(d['step2']['value'], 1), d['step2']['segments_ms'])
print('config5', d['config5'])
PY
