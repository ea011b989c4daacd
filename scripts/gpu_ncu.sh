ncu --set full --clock-control none --import-source on -k regex:"gather3" -s 4 -c 2 -o gpurun_out/gather -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
