#!/bin/bash
# usage: scripts/variants.sh <workload> <variant names...>; times the default build and every variant (one gpurun call)
wl=$1; shift
mkdir -p gpurun_out
python scripts/variant_bench.py $wl 2>&1 | tail -1
for v in "$@"; do
  LGM_LIB_PATH=$PWD/lagomorph_b200/variants/lib_$v.so python scripts/variant_bench.py $wl 2>&1 | tail -1
done
