(timeout 900 python -m pytest tests -m gpu -x -q -k "host or expmap" 2>&1 | tail -5) | tee gpurun_out/pytest_gpu.log
python scripts/e2e_sweep.py 2>&1 | grep -v Warning | tee gpurun_out/e2e_sweep.log
