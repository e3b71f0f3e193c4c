#!/bin/bash
# round 2, pass j (1 GPU): folded inference epilogue on the hot path (MODE 3) -- tests, A/B bench (forward-only + step 2)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_kernels.py tests/test_deeplab_gpu.py tests/test_step2_gpu.py tests/test_deeplab_parity.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r02j_tests.log 2>&1
echo "== tests exit $?"; tail -n 8 gpurun_out/r02j_tests.log
FLAGS="--steps 10 --warmup 3 --no-config5 --no-library-baseline --no-parity --no-numerics --no-cpu-baseline --no-transforms"
for hot in 0 1 0 1; do
  ZS3_FOLD_HOT=$hot timeout 600 python bench.py --gpus 1 $FLAGS > gpurun_out/r02j_bench_hot$hot.json 2> gpurun_out/r02j_bench_hot$hot.err
  python -c "
import json; d=json.load(open('gpurun_out/r02j_bench_hot$hot.json')); f=d['forward_only']; s=d['step2']
print('hot=$hot value', round(d['value'],1), 'fwd train graph', round(f['train_mode_bn_cuda_graph']['ms'],3), 'eval eager', round(f['eval_mode_bn_fused_epilogue']['ms'],3), 'eval graph', f.get('eval_mode_bn_fused_epilogue_cuda_graph'), 'step2', round(s['value'],1), s['segments_ms'])" || tail -n 5 gpurun_out/r02j_bench_hot$hot.err
done
