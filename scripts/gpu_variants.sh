for v in "" oldfluid; do
  if [ -n "$v" ]; then export LGM_LIB_PATH=$PWD/lagomorph_b200/variants/lib_$v.so; else unset LGM_LIB_PATH; fi
  python scripts/sharp_bench.py 16 128; python scripts/sharp_bench.py 4 256
done 2>&1 | grep -v Warning | tee gpurun_out/variants.log
