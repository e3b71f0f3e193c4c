#!/bin/bash
# A/B of ZS3_FPROP_SPLIT_N (128-column tiles for single-wave 256-channel convs) + step-2 segment profile.
mkdir -p gpurun_out
for v in 0 1; do
  ZS3_FPROP_SPLIT_N=$v timeout 200 python tools/conv_bench.py 30 > gpurun_out/conv_bench_split$v.md 2>&1
  echo "== conv_bench split_n=$v exit $?"; head -n 9 gpurun_out/conv_bench_split$v.md
done
ZS3_FPROP_SPLIT_N=1 timeout 300 python -m pytest tests/test_conv_kernels.py tests/test_deeplab_gpu.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/tests_split1.log 2>&1
echo "== conv/deeplab tests with split_n=1 exit $?"; tail -n 4 gpurun_out/tests_split1.log
for v in 0 1 0 1; do
  ZS3_FPROP_SPLIT_N=$v timeout 300 python bench.py --no-cpu-baseline --no-step2 --steps 20 > gpurun_out/bench_split${v}_$RANDOM.json 2> /dev/null
  echo "== bench split_n=$v exit $?"
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/bench_split*_*.json")):
    d = json.load(open(f))
    print(f, round(d["value"], 1), "img/s", round(d["ms_per_step"], 3), "ms; fwd train ms", round(d["forward_only"]["train_mode_bn"]["ms"], 3),
          "conv frac", round(d["roofline"]["frac"], 4), d["clocks"])
PY
timeout 300 python tools/step2_bench.py --steps 4 --warmup 2 --skip-unfused --out gpurun_out/step2_bench.json > gpurun_out/step2_bench.log 2>&1
echo "== step2 bench exit $?"; tail -n 3 gpurun_out/step2_bench.log
