#!/bin/bash
# ncu --set full of one launch of each quarter-slab kernel at C3 (8 x 256^3), text summary + source hot spots
mkdir -p gpurun_out
K='regex:xpassq|qslab'
ncu --set full --clock-control none --import-source on -k "$K" -s 3 -c 3 -o /tmp/r3_q_full -f python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/ncu_q.log 2>&1
python scripts/profile_summary.py /tmp/r3_q_full.ncu-rep gpurun_out/r3_ncu_qslab_summary.txt gpurun_out/r3_traffic_q.json "C3 share: 8 x 256^3" "ncu --set full --clock-control none --import-source on -k '$K' -s 3 -c 3 python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-extra"
ncu -i /tmp/r3_q_full.ncu-rep --page source --csv --kernel-name regex:xpassq > gpurun_out/r3_xpassq_source.csv 2>/dev/null
ls -la /tmp/*.ncu-rep gpurun_out/r3_*; tail -3 gpurun_out/ncu_q.log
