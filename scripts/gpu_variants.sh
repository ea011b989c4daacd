for v in "" h3; do
  if [ -n "$v" ]; then export LGM_LIB_PATH=$PWD/lagomorph_b200/variants/lib_$v.so; else unset LGM_LIB_PATH; fi
  python scripts/variant_bench.py c2; python scripts/variant_bench.py c3
done 2>&1 | grep -v Warning | tee gpurun_out/variants.log
