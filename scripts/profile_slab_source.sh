mkdir -p gpurun_out
K='regex:qslab_fwd|qslab_inv'
ncu --set full --clock-control none --import-source on -k "$K" -s 2 -c 2 -o /tmp/slabq -f python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-extra > /dev/null 2>&1
ncu -i /tmp/slabq.ncu-rep --page source --csv --kernel-name regex:qslab_fwd > gpurun_out/r3_qslab_fwd_source.csv 2>/dev/null
ncu -i /tmp/slabq.ncu-rep --page source --csv --kernel-name regex:qslab_inv > gpurun_out/r3_qslab_inv_source.csv 2>/dev/null
K='regex:slab_fwd_kernel|slab_inv_kernel'
ncu --set full --clock-control none --import-source on -k "$K" -s 2 -c 2 -o /tmp/slab128 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > /dev/null 2>&1
ncu -i /tmp/slab128.ncu-rep --page source --csv --kernel-name regex:slab_fwd_kernel > gpurun_out/r3_slab_fwd_source.csv 2>/dev/null
ncu -i /tmp/slab128.ncu-rep --page source --csv --kernel-name regex:slab_inv_kernel > gpurun_out/r3_slab_inv_source.csv 2>/dev/null
python __graft_entry__.py --smoke 2>&1 | tail -1
ls -la gpurun_out/r3_*source.csv
