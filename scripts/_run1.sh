python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "expmap_host" 2>&1 | tail -3
python scripts/clock_probe.py > gpurun_out/r3_clock_probe.log 2>&1
python bench.py --no-extra > gpurun_out/r3_bench_c2.json 2> gpurun_out/r3_bench_c2.err
tail -c 600 gpurun_out/r3_bench_c2.err
