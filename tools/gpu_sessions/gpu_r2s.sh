#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/profile_config5_host.py > gpurun_out/r02s_config5_host_profile.txt 2>&1
grep -A70 "cumulative" gpurun_out/r02s_config5_host_profile.txt | cut -c1-150 | head -75
