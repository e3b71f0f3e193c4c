#!/bin/bash
# round 2, first GPU pass: new split-precision kernels/engine tests, fused CE at 60 classes, then a short bench with all legs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in tests/test_parity_train_gpu.py tests/test_misc_kernels.py; do
  b=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -q -m gpu -s --tb=short -p no:cacheprovider > "gpurun_out/r02_${b}.log" 2>&1
  echo "== $f exit $?"
  tail -n 25 "gpurun_out/r02_${b}.log"
done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
echo "== bench exit $?"; tail -n 5 gpurun_out/r02_bench_a.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r02_bench_a.json'))
    print('value', d['value'], 'e2e', d['e2e']['value'])
    print('library', json.dumps(d.get('library_baseline'), indent=1)[:1500])
    print('parity', json.dumps(d.get('parity_mode'), indent=1)[:2500])
    print('numerics', json.dumps(d.get('numerics_vs_fp64'), indent=1)[:2500])
except Exception as e:
    print('no bench json', e)
PY
