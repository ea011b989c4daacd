#!/bin/bash
# quarter-slab path: correctness + timing against the cluster slab kernels (one gpurun call)
mkdir -p gpurun_out
python scripts/qslab_check.py 64 128 256 > gpurun_out/r3_qslab.log 2>&1
LGM_NO_QSLAB=1 python scripts/qslab_check.py 256 >> gpurun_out/r3_qslab.log 2>&1
python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r3_bench_c3_qslab.json 2> gpurun_out/r3_bench_c3_qslab.err
LGM_NO_QSLAB=1 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r3_bench_c3_cluster.json 2> gpurun_out/r3_bench_c3_cluster.err
tail -40 gpurun_out/r3_qslab.log
