#!/bin/bash
# usage: scripts/build_variant.sh <name> "<-D flags>" file1.cu [file2.cu ...]
# Builds lagomorph_b200/variants/lib_<name>.so: the listed csrc files recompiled with the flags, every
# other object taken from the main build (lagomorph_b200/build). Kernel experiments only
# (select with LGM_LIB_PATH); variants/ is git-ignored.
set -e
name=$1; flags=$2; shift; shift
root=$(cd $(dirname $0)/.. && pwd)
out=$root/lagomorph_b200/variants; mkdir -p $out/obj_$name
objs=""
for o in $root/lagomorph_b200/build/*.o; do
  b=$(basename $o .o); use=$o
  for f in "$@"; do
    if [ "$f" == "$b.cu" ]; then
      use=$out/obj_$name/$b.o
      /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $flags -c $root/lagomorph_b200/csrc/$f -o $use &
    fi
  done
  objs="$objs $use"
done
wait
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $out/lib_$name.so $objs
echo $out/lib_$name.so
