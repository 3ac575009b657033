#!/bin/bash
# build libb2enc.so variants with different K1 strip widths / residency (HERE, cross-compiled) into build_variants/
set -e
cd "$(dirname "$0")/../video-encoder_b200"
mkdir -p ../build_variants
for v in ${B2_K1_VARIANTS:-"8 2" "6 2" "5 3" "4 3" "4 4"}; do
  set -- ${v//_/ }
  rm -f csrc/k1_me_fullpel.o
  make -s NVCCFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-O2 -Xptxas -v -DB2_K1_NMB=$1 -DB2_K1_MINCTAS=$2 -DB2_K1_SMEM_PAD=${3:-0}" libb2enc.so
  cp libb2enc.so ../build_variants/libb2enc_nmb$1_c$2_pad${3:-0}.so
  grep -A3 "Li32ELi256ELb0" csrc/k1_me_fullpel.o.ptxas.log | grep -E "registers|spill" | head -2
done
rm -f csrc/k1_me_fullpel.o; make -s
