#!/bin/bash
# round 2, pass t (1 GPU): raw stream accessor + pinned plan uploads -- step-2 / GCN / parity tests, bench with all eager blocks
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_step2_gpu.py tests/test_gmmn_gpu.py tests/test_gmmn_fused_gpu.py tests/test_parity_train_gpu.py tests/test_transforms_gpu.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r02t_tests.log 2>&1
echo "== tests exit $?"; tail -n 4 gpurun_out/r02t_tests.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-library-baseline --no-numerics --no-cpu-baseline > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02t_bench.json'))
print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1))
s = d['step2']
print('step2', round(s['value'], 1), s['segments_ms'], 'e2e table', round(s['e2e_label_table_api']['value'], 1))
print('unchanged loop', s.get('unchanged_trainer_loop'))
print('config5', d['config5']['value'], d['config5']['segments_ms'])
print('parity', {k: (round(v['value'], 1), round(v['ms_per_step'], 2)) for k, v in d['parity_mode'].items()})
print('transforms', round(d['input_transforms']['value']), round(d['input_transforms']['e2e']['value']))
print('forward eager', {k: round(v['ms'], 3) for k, v in d['forward_only'].items()})
PY
