#!/bin/bash
# round 2, 2-GPU pass: multi-GPU hardware tests, then 2-rank bench (step 1 + step 2 + config 5)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -s --tb=short -p no:cacheprovider > gpurun_out/r02c_multigpu.log 2>&1
echo "== multigpu tests exit $?"; grep -E "rel-L2|passed|failed|Error|error" gpurun_out/r02c_multigpu.log | tail -n 20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02c_bench_dp2.json 2> gpurun_out/r02c_bench_dp2.err
echo "== bench dp2 exit $?"; tail -n 4 gpurun_out/r02c_bench_dp2.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r02c_bench_dp2.json'))
    print('value', d['value'], 'e2e', d['e2e']['value'])
    print('step2', json.dumps(d.get('step2'))[:600])
    print('config5', json.dumps(d.get('config5'))[:600])
except Exception as e:
    print('no bench json', e)
PY
