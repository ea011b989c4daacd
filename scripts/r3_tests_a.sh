#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_shoot_gpu.py tests/test_epdiff_bwd_gpu.py tests/test_fullsize_gpu.py tests/test_fullsize_ref_gpu.py tests/test_atlas_oracle_gpu.py -x -q -m gpu > gpurun_out/r3_pytest_a.log 2>&1
tail -5 gpurun_out/r3_pytest_a.log
python scripts/variant_bench.py c2 2>&1 | tail -1
python scripts/variant_bench.py c3 2>&1 | tail -1
LGM_NO_FIRST_STEP_SHORTCUT=1 python scripts/variant_bench.py c3 2>&1 | tail -1
python scripts/atlas_profile.py 2>&1 | head -12
