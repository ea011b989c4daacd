python scripts/variant_bench.py c2 2>&1 | grep -v Warning | tee gpurun_out/variants.log
LGM_NO_ALTERNATE=1 python scripts/variant_bench.py c2 2>&1 | grep -v Warning | tee -a gpurun_out/variants.log
python scripts/variant_bench.py c3 2>&1 | grep -v Warning | tee -a gpurun_out/variants.log
LGM_NO_ALTERNATE=1 python scripts/variant_bench.py c3 2>&1 | grep -v Warning | tee -a gpurun_out/variants.log
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) | tee gpurun_out/pytest_gpu.log
