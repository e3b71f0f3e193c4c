#!/bin/bash
# round 2, second GPU pass: reworked parity tests, step-2 golden / table / GCN tests, bench with the new step2 block
mkdir -p gpurun_out
for f in tests/test_parity_train_gpu.py tests/test_step2_gpu.py; do
  b=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -q -m gpu -s --tb=short -p no:cacheprovider > "gpurun_out/r02b_${b}.log" 2>&1
  echo "== $f exit $?"
  grep -E "pieces=|train-BN|C=|513|  grad |bf16 path|g_losses|passed|failed|Error" "gpurun_out/r02b_${b}.log" | tail -n 60
done
timeout 1200 python bench.py --steps 5 --warmup 3 --no-library-baseline --no-parity --no-numerics > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
echo "== bench exit $?"; tail -n 3 gpurun_out/r02b_bench.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r02b_bench.json'))
    print('value', d['value'], 'e2e', d['e2e']['value'])
    print('step2', json.dumps(d.get('step2'), indent=1)[:4000])
except Exception as e:
    print('no bench json', e)
PY
