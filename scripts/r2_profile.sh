#!/bin/bash
# (the .ncu-rep files stay on the box: with --import-source they exceed gpurun's 64 MiB return limit)
# round-2 captures on one B200: ncu launch list of the bench command, ncu --set full of one launch of
# each EPDiff-step kernel at C2 and C3, text summaries + DRAM traffic (profiles/r2_*)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu_bench.log 2>&1
K='regex:gather3|xpass2|slab|compose_ring'
ncu --set full --clock-control none --import-source on -k "$K" -s 5 -c 5 -o /tmp/r2_c2_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" -s 5 -c 5 -o /tmp/r2_c3_full -f python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/ncu_c3.log 2>&1
python scripts/profile_summary.py /tmp/r2_c2_full.ncu-rep gpurun_out/r2_ncu_full_summary.txt gpurun_out/r2_traffic.json "C2: 16 x 128^3" "ncu --set full --clock-control none --import-source on -k '$K' -s 5 -c 5 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra"
python scripts/profile_summary.py /tmp/r2_c3_full.ncu-rep gpurun_out/r2_ncu_full_summary_c3.txt gpurun_out/r2_traffic_c3.json "C3 share: 8 x 256^3" "ncu --set full ... python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-extra"
ls -la /tmp/*.ncu-rep; rm -f gpurun_out/ncu_*.log
