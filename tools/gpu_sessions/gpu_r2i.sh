#!/bin/bash
# round 2, pass i (8 GPUs): the driver's scaling command at N=8 and N=4 with the two-stage backward (default) -- full default blocks at N=8 once
mkdir -p gpurun_out
FLAGS="--steps 10 --warmup 3 --no-library-baseline --no-parity --no-numerics --no-cpu-baseline"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 $FLAGS > gpurun_out/r02i_bench_dp8.json 2> gpurun_out/r02i_bench_dp8.err
echo "== bench dp8 exit $?"; tail -n 3 gpurun_out/r02i_bench_dp8.err
ZS3_DP_CUT=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 $FLAGS --no-step2 --no-config5 > gpurun_out/r02i_bench_dp8_cut0.json 2> gpurun_out/r02i_bench_dp8_cut0.err
echo "== bench dp8 cut0 exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 $FLAGS --no-step2 --no-config5 > gpurun_out/r02i_bench_dp4.json 2> gpurun_out/r02i_bench_dp4.err
echo "== bench dp4 exit $?"
timeout 600 python bench.py --gpus 1 $FLAGS --no-step2 --no-config5 > gpurun_out/r02i_bench_dp1.json 2> gpurun_out/r02i_bench_dp1.err
python - <<'PY'
import json
for f in ('dp8', 'dp8_cut0', 'dp4', 'dp1'):
    try:
        d = json.load(open(f'gpurun_out/r02i_bench_{f}.json'))
        print(f, 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), 'loss', d['final_loss'],
              'step2', (d.get('step2') or {}).get('value'), 'config5', (d.get('config5') or {}).get('value'))
    except Exception as e:
        print(f, 'failed', e)
PY
