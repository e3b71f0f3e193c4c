#!/bin/bash
# round 2, pass q (1 GPU): the driver's default bench (all blocks) + reference arm
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err
echo "== bench exit $?"; tail -n 2 gpurun_out/r02q_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02q_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'roofline frac', d['roofline']['frac'])
print('forward_only', {k: round(v['ms'], 3) for k, v in d['forward_only'].items()})
print('step2', d['step2'].get('value'), d['step2'].get('segments_ms'), 'e2e table', d['step2'].get('e2e_label_table_api', {}).get('value'))
print('config5', d['config5'])
print('transforms', d['input_transforms']['value'], d['input_transforms']['e2e']['value'])
print('library', {k: (v.get('value') if isinstance(v, dict) else None) for k, v in d['library_baseline'].items()})
print('parity', {k: v.get('value') for k, v in d['parity_mode'].items()})
print('cpu', d['cpu_baseline'])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02q_bench_reference.json 2> gpurun_out/r02q_bench_reference.err
echo "== reference arm exit $?"; cat gpurun_out/r02q_bench_reference.json | head -c 1200
