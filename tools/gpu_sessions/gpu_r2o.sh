#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ce_bench.py 21 20
timeout 300 python tools/ce_bench.py 60 10
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"upsample4_ce_bwd_kernel|upsample_ce_fwd" --launch-count 4 -o gpurun_out/r02o_ce_full -f python tools/ce_bench.py 21 1 > gpurun_out/r02o_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r02o_ce_full.ncu-rep
ncu -i gpurun_out/r02o_ce_full.ncu-rep --page details --csv 2>/dev/null | grep -E "upsample4" | grep -iE "Stall|Issue|Eligible|Bank|Achieved Occupancy|Registers|Theoretical Occ|Executed Ipc|No Eligible" | awk -F'","' '{print $(NF-5), $(NF-3), $(NF-2), $(NF-1)}' | sort -u | head -40
