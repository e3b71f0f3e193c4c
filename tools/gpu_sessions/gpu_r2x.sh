#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 tools/scale_probe.py 20 > gpurun_out/r02x_scale_probe_2.json 2> gpurun_out/r02x_scale_probe_2.err
echo "== probe exit $?"; cat gpurun_out/r02x_scale_probe_2.json; tail -n 2 gpurun_out/r02x_scale_probe_2.err
