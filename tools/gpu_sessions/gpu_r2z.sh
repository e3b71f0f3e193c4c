#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02z_bench_dp8.json 2> gpurun_out/r02z_bench_dp8.err
echo "== bench dp8 exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r02z_bench_dp8.json')); print('dp8 value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'loss', d['final_loss'], 'step2', round((d.get('step2') or {}).get('value',0),1), 'config5', round((d.get('config5') or {}).get('value',0),1))" || tail -n 8 gpurun_out/r02z_bench_dp8.err
