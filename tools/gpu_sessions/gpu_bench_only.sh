#!/bin/bash
mkdir -p gpurun_out
timeout 230 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "== bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print(d['value'], d['e2e']['value'], d['forward_only']['train_mode_bn_cuda_graph'], d['step2']['value'], d['roofline']['frac'], d['cpu_baseline']['value'])"
