#!/bin/bash
# round 2, pass f (2 GPUs): two-stage backward with overlapped all-reduce -- hardware test, then A/B bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -s --tb=short -p no:cacheprovider -k two_stage > gpurun_out/r02f_dp_cut_test.log 2>&1
echo "== cut test exit $?"; grep -E "rel-L2|passed|failed|Error|error|assert" gpurun_out/r02f_dp_cut_test.log | tail -n 20
FLAGS="--gpus 2 --steps 10 --warmup 3 --no-step2 --no-config5 --no-library-baseline --no-parity --no-numerics --no-cpu-baseline"
for cut in 0 1 0 1; do
  ZS3_DP_CUT=$cut timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$cut bench.py $FLAGS > gpurun_out/r02f_bench_dp2_cut$cut.json 2> gpurun_out/r02f_bench_dp2_cut$cut.err
  echo "== bench cut=$cut exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r02f_bench_dp2_cut$cut.json')); print('cut=$cut value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])" || tail -n 5 gpurun_out/r02f_bench_dp2_cut$cut.err
done
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-step2 --no-config5 --no-library-baseline --no-parity --no-numerics --no-cpu-baseline > gpurun_out/r02f_bench_dp1.json 2> gpurun_out/r02f_bench_dp1.err
python -c "
import json; d=json.load(open('gpurun_out/r02f_bench_dp1.json')); print('N=1 value', d['value'], 'ms', d['ms_per_step'])"
