ncu --set full --clock-control none --import-source on -k regex:"cslab|xpass2" -s 3 -c 3 -o gpurun_out/cslab -f python scripts/sharp_once.py 4 256 > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
