#!/bin/bash
# round 2, pass h (1 GPU): default bench (all blocks) with the per-conv-shape table; transforms tests again (stage/launch split)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_transforms_gpu.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -n 3
timeout 900 python bench.py --steps 10 --warmup 3 --layer-table gpurun_out/r02h_layer_table.md > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
echo "== bench exit $?"; tail -n 3 gpurun_out/r02h_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02h_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
print('transforms', json.dumps(d.get('input_transforms'))[:1500])
print('step2', d['step2'].get('value'), 'config5', d['config5'].get('value'))
PY
head -n 45 gpurun_out/r02h_layer_table.md
