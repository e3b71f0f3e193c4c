#!/bin/bash
# round 2, pass w (2 GPUs): final validation of the build -- GPU suite in the driver's order, smoke, default bench (N=1), 2-GPU bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --tb=short --durations=8 -p no:cacheprovider > gpurun_out/r02w_gpu_suite.log 2>&1
echo "== gpu suite exit $?"; tail -n 14 gpurun_out/r02w_gpu_suite.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02w_bench.json 2> gpurun_out/r02w_bench.err
echo "== bench exit $?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/r02w_bench.json'))
print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), 'frac', round(d['roofline']['frac'], 3))
s = d['step2']
print('step2', round(s['value'], 1), s['segments_ms'], 'e2e table', round(s['e2e_label_table_api']['value'], 1), 'unchanged', round(s['unchanged_trainer_loop']['value'], 1), 'lib', round(s['library_baseline']['value'], 1))
print('config5', round(d['config5']['value'], 1), 'transforms', round(d['input_transforms']['value']), round(d['input_transforms']['e2e']['value']))
print('library', {k: round(v['value'], 1) for k, v in d['library_baseline'].items() if isinstance(v, dict)}, 'parity', {k: round(v['value'], 1) for k, v in d['parity_mode'].items()})
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02w_bench_dp2.json 2> gpurun_out/r02w_bench_dp2.err
echo "== bench dp2 exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r02w_bench_dp2.json')); print('dp2 value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'step2', round((d.get('step2') or {}).get('value',0),1), 'config5', round((d.get('config5') or {}).get('value',0),1))"
