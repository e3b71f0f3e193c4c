#!/bin/bash
# round 2, pass r (2 GPUs): the whole GPU suite in the driver's order on a 2-GPU box + smoke + 2-GPU bench (driver's launch line)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r02r_gpu_suite.log 2>&1
echo "== gpu suite exit $?"; tail -n 6 gpurun_out/r02r_gpu_suite.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02r_bench_dp2.json 2> gpurun_out/r02r_bench_dp2.err
echo "== bench dp2 exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r02r_bench_dp2.json')); print('dp2 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'step2', (d.get('step2') or {}).get('value'), 'config5', (d.get('config5') or {}).get('value'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02r_bench_ref_dp2.json 2> gpurun_out/r02r_bench_ref_dp2.err
echo "== reference arm dp2 exit $?"; head -c 300 gpurun_out/r02r_bench_ref_dp2.json
