#!/bin/bash
# full GPU suite in the driver's order (single GPU: the 2-GPU tests skip)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r02k_gpu_suite.log 2>&1
echo "== gpu suite exit $?"; tail -n 8 gpurun_out/r02k_gpu_suite.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
