#!/bin/bash
# round 2, pass p (1 GPU): stem im2col staging + maxpool backward staging -- tests, kernel times, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_misc_kernels.py tests/test_deeplab_gpu.py tests/test_parity_train_gpu.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r02p_tests.log 2>&1
echo "== tests exit $?"; tail -n 4 gpurun_out/r02p_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -k regex:"maxpool|stem_im2col|upsample|bilinear" --csv --log-file gpurun_out/r02p_glue_kernels.csv python tools/profile_step.py 16 1 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02p_glue_kernels.csv
FLAGS="--steps 10 --warmup 3 --no-step2 --no-config5 --no-library-baseline --no-parity --no-numerics --no-cpu-baseline --no-transforms"
for i in 1 2; do
  timeout 600 python bench.py --gpus 1 $FLAGS > gpurun_out/r02p_bench_$i.json 2> gpurun_out/r02p_bench_$i.err
  python -c "
import json; d=json.load(open('gpurun_out/r02p_bench_$i.json'))
print('run $i value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'loss', d['final_loss'], 'fwd graph', round(d['forward_only']['train_mode_bn_cuda_graph']['ms'],3))" || tail -n 5 gpurun_out/r02p_bench_$i.err
done
