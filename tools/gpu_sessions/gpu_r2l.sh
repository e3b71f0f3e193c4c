#!/bin/bash
# round 2, pass l (1 GPU): wgrad tile choice by cost model -- tests, A/B bench with the per-shape table
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_kernels.py tests/test_deeplab_gpu.py tests/test_parity_train_gpu.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r02l_tests.log 2>&1
echo "== tests exit $?"; tail -n 4 gpurun_out/r02l_tests.log
FLAGS="--steps 10 --warmup 3 --no-step2 --no-config5 --no-library-baseline --no-parity --no-numerics --no-cpu-baseline --no-transforms"
for tile in 2x256 auto 2x256 auto; do
  if [ "$tile" = "auto" ]; then unset ZS3_WGRAD_TILE; else export ZS3_WGRAD_TILE=$tile; fi
  timeout 600 python bench.py --gpus 1 $FLAGS --layer-table gpurun_out/r02l_layer_table_$tile.md > gpurun_out/r02l_bench_$tile.json 2> gpurun_out/r02l_bench_$tile.err
  python -c "
import json; d=json.load(open('gpurun_out/r02l_bench_$tile.json')); r=d['roofline']
print('tile=$tile value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'wgrad ms', round(r['by_kind']['conv_wgrad']['ms_per_step'],3), 'conv ms', round(r['conv_ms_per_step'],3), 'loss', d['final_loss'])" || tail -n 5 gpurun_out/r02l_bench_$tile.err
done
grep "conv_wgrad" gpurun_out/r02l_layer_table_auto.md | head -12
echo; grep "conv_wgrad" gpurun_out/r02l_layer_table_2x256.md | head -12
