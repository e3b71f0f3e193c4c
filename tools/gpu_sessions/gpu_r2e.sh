#!/bin/bash
# round 2, pass e (1 GPU): device transforms -- parity tests, timing, launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_transforms_gpu.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r02e_test_transforms_gpu.log 2>&1
echo "== transforms tests exit $?"; tail -n 25 gpurun_out/r02e_test_transforms_gpu.log
timeout 300 python tools/augment_bench.py 16 20 > gpurun_out/r02e_augment_bench.json 2> gpurun_out/r02e_augment_bench.err
echo "== augment bench exit $?"; cat gpurun_out/r02e_augment_bench.json; tail -n 5 gpurun_out/r02e_augment_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:aug_ -c 40 --csv --log-file gpurun_out/r02e_augment_launches.csv python tools/augment_bench.py 16 2 > /dev/null 2>&1
echo "== ncu exit $?"; tail -n 12 gpurun_out/r02e_augment_launches.csv
