#!/bin/bash
# ncu --set full of one launch of each backward kernel of the fused EPDiff step at 8 x 256^3 (atlas epoch)
mkdir -p gpurun_out
K='regex:adstar_bwd3|compose_bwd3|stencil_bwd3'
ncu --set full --clock-control none --import-source on -k "$K" -s 6 -c 3 -o /tmp/bwd_full -f python scripts/atlas_profile.py > gpurun_out/ncu_bwd.log 2>&1
python scripts/profile_summary.py /tmp/bwd_full.ncu-rep gpurun_out/r2_ncu_bwd_summary.txt gpurun_out/r2_traffic_bwd.json "atlas epoch, 8 x 256^3" "ncu --set full --clock-control none --import-source on -k '$K' -s 6 -c 3 python scripts/atlas_profile.py"
ncu -i /tmp/bwd_full.ncu-rep --page raw --csv > gpurun_out/r2_bwd_raw.csv 2>/dev/null
tail -3 gpurun_out/ncu_bwd.log
