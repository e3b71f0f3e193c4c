#!/bin/bash
# round 2, pass n (1 GPU): fused upsample+CE backward without per-row block syncs -- tests, kernel time, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_misc_kernels.py tests/test_deeplab_gpu.py tests/test_step2_gpu.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r02n_tests.log 2>&1
echo "== tests exit $?"; tail -n 4 gpurun_out/r02n_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:upsample_ce --csv --log-file gpurun_out/r02n_ce_kernels.csv python tools/profile_step.py 16 1 > /dev/null 2>&1
grep -E "upsample_ce" gpurun_out/r02n_ce_kernels.csv | cut -d, -f5,13-15
FLAGS="--steps 10 --warmup 3 --no-config5 --no-library-baseline --no-parity --no-numerics --no-cpu-baseline --no-transforms"
for i in 1 2; do
  timeout 600 python bench.py --gpus 1 $FLAGS > gpurun_out/r02n_bench_$i.json 2> gpurun_out/r02n_bench_$i.err
  python -c "
import json; d=json.load(open('gpurun_out/r02n_bench_$i.json')); s=d['step2']
print('run $i value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'loss', d['final_loss'], 'step2', round(s['value'],1), s['segments_ms'])" || tail -n 5 gpurun_out/r02n_bench_$i.err
done
