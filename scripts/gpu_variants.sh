for v in "" nr2 nr4; do
  if [ -n "$v" ]; then export LGM_LIB_PATH=$PWD/lagomorph_b200/variants/lib_$v.so; else unset LGM_LIB_PATH; fi
  python scripts/variant_bench.py c2
done 2>&1 | grep -v Warning | tee gpurun_out/variants.log
LGM_LIB_PATH=$PWD/lagomorph_b200/variants/lib_nr2.so timeout 600 python -m pytest tests -m gpu -x -q -k "adjrep or fullsize or golden or expmap" 2>&1 | tail -2
