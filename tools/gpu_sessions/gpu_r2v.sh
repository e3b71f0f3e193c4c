#!/bin/bash
# round 2, pass v (1 GPU): weight-gradient kernels on a side stream (ZS3_WGRAD_STREAM) -- tests, A/B bench
mkdir -p gpurun_out
ZS3_WGRAD_STREAM=1 timeout 900 python -m pytest tests/test_deeplab_gpu.py tests/test_parity_train_gpu.py tests/test_conv_kernels.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r02v_tests.log 2>&1
echo "== tests (side stream on) exit $?"; tail -n 4 gpurun_out/r02v_tests.log
FLAGS="--steps 10 --warmup 3 --no-step2 --no-config5 --no-library-baseline --no-parity --no-numerics --no-cpu-baseline --no-transforms"
for ws in 0 1 0 1; do
  ZS3_WGRAD_STREAM=$ws timeout 600 python bench.py --gpus 1 $FLAGS > gpurun_out/r02v_bench_ws$ws.json 2> gpurun_out/r02v_bench_ws$ws.err
  python -c "
import json; d=json.load(open('gpurun_out/r02v_bench_ws$ws.json'))
print('wgrad_stream=$ws value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'loss', d['final_loss'])" || tail -n 5 gpurun_out/r02v_bench_ws$ws.err
done
