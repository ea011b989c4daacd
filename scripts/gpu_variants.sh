for v in "" xm3 xm5 tx256; do
  if [ -n "$v" ]; then export LGM_LIB_PATH=$PWD/lagomorph_b200/variants/lib_$v.so; else unset LGM_LIB_PATH; fi
  python scripts/sharp_bench.py 16 128 | head -1; python scripts/sharp_bench.py 8 256 | head -1
done 2>&1 | grep -v Warning | tee gpurun_out/variants.log
