#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gmmn_fused_gpu.py tests/test_step2_gpu.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_fused_v5.log 2>&1
echo "== fused/step2 tests exit $?"; tail -n 4 gpurun_out/test_fused_v5.log
timeout 200 python tools/ncu_new_kernels.py > gpurun_out/new_kernels_timing.json 2> gpurun_out/new_kernels_timing.err
echo "== timings exit $?"; cat gpurun_out/new_kernels_timing.json
ZS3_NCU=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gmmn_train_fused|label_components|argmax_confusion" \
  --launch-count 6 -o gpurun_out/new_kernels_full -f python tools/ncu_new_kernels.py > gpurun_out/ncu_new.log 2>&1
echo "== ncu exit $?"; tail -n 2 gpurun_out/ncu_new.log
python tools/ncu_summary.py gpurun_out/new_kernels_full.ncu-rep > gpurun_out/new_kernels_ncu_summary.md 2>&1; cat gpurun_out/new_kernels_ncu_summary.md
