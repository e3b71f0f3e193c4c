#!/bin/bash
# One GPU session: full GPU test suite (one process per file), step-2 bench, new-kernel timings + ncu, default bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in tests/test_gmmn_fused_gpu.py tests/test_graph_gpu.py tests/test_metrics_gpu.py tests/test_step2_gpu.py; do
  b=$(basename "$f" .py)
  timeout 420 python -m pytest "$f" -q -m gpu --tb=short -p no:cacheprovider > "gpurun_out/${b}.log" 2>&1
  echo "== $f exit $?"; tail -n 15 "gpurun_out/${b}.log"
done
timeout 240 python tools/ncu_new_kernels.py > gpurun_out/new_kernels_timing.json 2> gpurun_out/new_kernels_timing.err
echo "== new kernel timings exit $?"; cat gpurun_out/new_kernels_timing.json; tail -n 5 gpurun_out/new_kernels_timing.err
timeout 420 python tools/step2_bench.py --steps 4 --warmup 2 --out gpurun_out/step2_bench.json > gpurun_out/step2_bench.log 2>&1
echo "== step2 bench exit $?"; tail -n 5 gpurun_out/step2_bench.log
ZS3_NCU=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gmmn_train_fused|label_components|argmax_confusion" \
  --launch-count 6 -o gpurun_out/new_kernels_full -f python tools/ncu_new_kernels.py > gpurun_out/ncu_new.log 2>&1
echo "== ncu exit $?"; tail -n 3 gpurun_out/ncu_new.log
python tools/ncu_summary.py gpurun_out/new_kernels_full.ncu-rep > gpurun_out/new_kernels_ncu_summary.md 2>&1; cat gpurun_out/new_kernels_ncu_summary.md
timeout 600 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider --deselect tests/test_gmmn_fused_gpu.py \
  --deselect tests/test_graph_gpu.py --deselect tests/test_metrics_gpu.py --deselect tests/test_step2_gpu.py > gpurun_out/gpu_tests_rest.log 2>&1
echo "== rest of gpu suite exit $?"; tail -n 6 gpurun_out/gpu_tests_rest.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "== bench exit $?"; cat gpurun_out/bench_default.json | cut -c1-600
