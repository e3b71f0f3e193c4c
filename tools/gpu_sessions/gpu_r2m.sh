#!/bin/bash
# round 2, pass m (1 GPU): ncu evidence of the round-2 build -- launch list of one training step, full capture of the
# conv / BN shapes, full capture of the input-transform kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02m_launches_step.csv python tools/profile_step.py 16 1 > gpurun_out/r02m_profile_step.log 2>&1
echo "== launch list exit $?"
python tools/summarize_launches.py gpurun_out/r02m_launches_step.csv > gpurun_out/r02m_launches_step.md 2>&1; head -n 24 gpurun_out/r02m_launches_step.md
python tools/conv_traffic.py gpurun_out/r02m_launches_step.csv gpurun_out/r02m_conv_traffic.json > /dev/null 2>&1; head -c 700 gpurun_out/r02m_conv_traffic.json; echo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_fprop|conv_wgrad" --launch-count 18 -o gpurun_out/r02m_conv_full -f python tools/ncu_shapes.py > gpurun_out/r02m_ncu_conv.log 2>&1
echo "== ncu conv exit $?"
python tools/ncu_summary.py gpurun_out/r02m_conv_full.ncu-rep > gpurun_out/r02m_ncu_conv_summary.md 2>&1; cat gpurun_out/r02m_ncu_conv_summary.md
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"aug_" --launch-count 10 -o gpurun_out/r02m_aug_full -f python tools/augment_bench.py 16 1 > gpurun_out/r02m_ncu_aug.log 2>&1
echo "== ncu aug exit $?"
python tools/ncu_summary.py gpurun_out/r02m_aug_full.ncu-rep > gpurun_out/r02m_ncu_aug_summary.md 2>&1; cat gpurun_out/r02m_ncu_aug_summary.md
ls -la gpurun_out/*.ncu-rep
