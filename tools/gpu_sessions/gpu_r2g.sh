#!/bin/bash
# round 2, pass g (2 GPUs): tap culling (conv tests, model tests, A/B bench on 1 GPU) + cut-backward/side-stream optimizer (2-GPU test, A/B bench)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_kernels.py tests/test_deeplab_gpu.py tests/test_parity_train_gpu.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r02g_conv_tests.log 2>&1
echo "== conv/model tests exit $?"; tail -n 6 gpurun_out/r02g_conv_tests.log
timeout 600 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -s --tb=short -p no:cacheprovider -k two_stage > gpurun_out/r02g_dp_cut_test.log 2>&1
echo "== cut test exit $?"; grep -E "rel-L2|passed|failed|Error|error|assert" gpurun_out/r02g_dp_cut_test.log | tail -n 20
FLAGS="--steps 10 --warmup 3 --no-step2 --no-config5 --no-library-baseline --no-parity --no-numerics --no-cpu-baseline"
for cull in 0 1 0 1; do
  ZS3_TAP_CULL=$cull timeout 600 python bench.py --gpus 1 $FLAGS > gpurun_out/r02g_bench_cull$cull.json 2> gpurun_out/r02g_bench_cull$cull.err
  python -c "
import json; d=json.load(open('gpurun_out/r02g_bench_cull$cull.json')); print('cull=$cull value', d['value'], 'ms', d['ms_per_step'], 'fwd graph ms', d['forward_only']['train_mode_bn_cuda_graph']['ms'], 'conv ms', d['roofline']['conv_ms_per_step'], 'loss', d['final_loss'])" || tail -n 5 gpurun_out/r02g_bench_cull$cull.err
done
for cut in 0 1 0 1; do
  ZS3_DP_CUT=$cut timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$cut bench.py --gpus 2 $FLAGS > gpurun_out/r02g_bench_dp2_cut$cut.json 2> gpurun_out/r02g_bench_dp2_cut$cut.err
  python -c "
import json; d=json.load(open('gpurun_out/r02g_bench_dp2_cut$cut.json')); print('cut=$cut value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'loss', d['final_loss'])" || tail -n 5 gpurun_out/r02g_bench_dp2_cut$cut.err
done
