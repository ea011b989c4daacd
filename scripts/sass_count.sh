#!/bin/bash
# usage: scripts/sass_count.sh <file.cu> [pattern] [extra nvcc flags]: SASS instruction count + registers per kernel
src=$1; pat=${2:-.}; shift; shift
obj=/tmp/sasscount_$$.o
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xptxas -v "$@" -c $src -o $obj 2> /tmp/sasscount_$$.log || { tail -20 /tmp/sasscount_$$.log; exit 1; }
cuobjdump -sass $obj | awk -v pat="$pat" '
/Function :/ { if (name != "" && name ~ pat) printf "%6d instr  LDG %3d LDS %3d STS %3d  %s\n", n, ldg, lds, sts, name; name=$3; n=0; ldg=0; lds=0; sts=0 }
/^[ \t]+\/\*[0-9a-f]+\*\/ / { n++; if ($0 ~ /LDG/) ldg++; if ($0 ~ /LDS/) lds++; if ($0 ~ /STS/) sts++ }
END { if (name ~ pat) printf "%6d instr  LDG %3d LDS %3d STS %3d  %s\n", n, ldg, lds, sts, name }'
grep -B1 -A3 "Compiling entry function" /tmp/sasscount_$$.log | grep -E "entry function|registers" | paste - - | grep -E "$pat" | sed -E "s/.*function '([^']*)'.*Used ([0-9]+) registers.*/\2 regs \1/"
rm -f $obj /tmp/sasscount_$$.log
