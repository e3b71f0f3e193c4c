#!/bin/bash
# final-state check: smoke(), whole GPU suite, default bench line
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "== smoke exit $?"; tail -n 3 gpurun_out/smoke.log
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/gpu_tests_all.log 2>&1
echo "== gpu suite exit $?"; tail -n 5 gpurun_out/gpu_tests_all.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "== bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print(d['value'], d['e2e']['value'], d['forward_only'], d['step2']['value'])"
