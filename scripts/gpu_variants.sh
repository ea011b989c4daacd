python scripts/e2e_sweep.py 2>&1 | grep -v Warning | grep expmap_host | tee gpurun_out/e2e_sweep.log
