#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02y_bench_dp4.json 2> gpurun_out/r02y_bench_dp4.err
echo "== bench dp4 exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r02y_bench_dp4.json')); print('dp4 value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'loss', d['final_loss'], 'step2', round((d.get('step2') or {}).get('value',0),1), 'config5', round((d.get('config5') or {}).get('value',0),1))" || tail -n 8 gpurun_out/r02y_bench_dp4.err
