#!/bin/bash
# round-2 experiment 1: packed shoot vs planar step loop
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_shoot_gpu.py -x -q) > gpurun_out/exp1_pytest.log 2>&1
tail -3 gpurun_out/exp1_pytest.log
for wl in c2 c3; do
  LGM_NO_PACKED=1 python scripts/variant_bench.py $wl
  python scripts/variant_bench.py $wl
  for v in a2 a4 nv2 nv1 c3 c5; do
    LGM_LIB_PATH=$PWD/lagomorph_b200/variants/lib_$v.so python scripts/variant_bench.py $wl
  done
done 2>&1 | tee gpurun_out/exp1_variants.log
