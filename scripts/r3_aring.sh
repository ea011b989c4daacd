#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_shoot_gpu.py tests/test_parity_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu -k "adstar or Ad_star or ad_star or ring or shoot or fullsize or coad" > gpurun_out/r3_pytest_b.log 2>&1
tail -4 gpurun_out/r3_pytest_b.log
for wl in c2 c3; do
python scripts/variant_bench.py $wl 2>&1 | tail -1
LGM_NO_ADSTAR_RING=1 python scripts/variant_bench.py $wl 2>&1 | tail -1
done
