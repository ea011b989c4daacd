#!/bin/bash
# SASS evidence of the shipped library: per kernel instruction counts and the memory / cluster / async
# mnemonics that matter (profiles/r2_sass_counts.txt). usage: scripts/sass_summary.sh > profiles/r2_sass_counts.txt
so=$(dirname $0)/../lagomorph_b200/liblagomorph_b200.so
echo "# cuobjdump -sass $(basename $so)  ($(date -u +%F))"
cuobjdump -sass $so > /tmp/sass_all.$$ 
echo "## whole library: mnemonic counts"
for m in LDG STG LDS STS REDG ATOMG RED ATOMS LDGSTS UTMALDG UTMASTG UBLKCP UBLKPF UCGABAR SYNCS FADD2 FFMA2 FMUL2 SHFL BAR.SYNC; do
  printf "%-8s %7d\n" $m $(grep -c "[ .]$m" /tmp/sass_all.$$)
done
echo "functions: $(grep -c 'Function :' /tmp/sass_all.$$)   arch: $(grep -m1 -o 'sm_[0-9a-z]*' /tmp/sass_all.$$)"
echo "## per kernel (EPDiff step and its backward, FFT passes, splats)"
awk '
/Function :/ { if (name != "") out(); name=$3; n=0; ldg=0; stg=0; lds=0; sts=0; red=0; shf=0; bar=0; cg=0; tma=0; f2=0 }
/^[ \t]+\/\*[0-9a-f]+\*\/ / { n++; if ($0 ~ /LDG/) ldg++; if ($0 ~ /STG/) stg++; if ($0 ~ /LDS/) lds++; if ($0 ~ /STS/) sts++;
  if ($0 ~ /RED|ATOMG/) red++; if ($0 ~ /SHFL/) shf++; if ($0 ~ /BAR\.SYNC/) bar++; if ($0 ~ /UCGABAR/) cg++; if ($0 ~ /UTMALDG|UBLKCP|UBLKPF|LDGSTS/) tma++; if ($0 ~ /FADD2|FFMA2|FMUL2/) f2++ }
function out() { if (name ~ /gather3_kernel|ring_kernel|slab_|xpass2|adstar_bwd|compose_bwd|stencil_bwd|splat3|interp3|interp_du3|affine3|zfwd|zinv|ypass2|ad_star3|jtvf/)
  printf "%6d instr LDG %3d STG %3d LDS %3d STS %3d RED %3d SHFL %3d BAR %2d CGA %2d ASYNC %2d F2 %3d  %s\n", n, ldg, stg, lds, sts, red, shf, bar, cg, tma, f2, name }
END { out() }' /tmp/sass_all.$$ | sort -k22 | cut -c1-230
rm -f /tmp/sass_all.$$
