#!/bin/sh
# EXPERIMENT: the extraction kernel with the exact-arithmetic constraint dropped -- MUFU square root and
# logarithm, contracted multiply-adds -- built beside the product library, never instead of it.
#   sh tools/relaxed_experiment.sh           builds tools/_build/libtiresias_gpu_relaxed.so   (here, no GPU needed)
#   python tools/gpu_relaxed_experiment.py   times both libraries and measures the gates      (on the GPU box)
set -e
cd "$(dirname "$0")/../asterisk_tiresias_b200/csrc"
OUT=../../tools/_build
mkdir -p $OUT
OBJS=""
for f in tir_api.cu tir_extract.cu tir_match.cu tir_p2p.cu tir_stream.cu tir_tables.cpp tir_batcher.cpp tir_sqlite.cpp tir_group.cpp; do
  o=$OUT/relaxed_$(basename $f | sed 's/\.[a-z]*$//').o
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DTIR_RELAXED -Xcompiler -fPIC,-Wall -c $f -o $o
  OBJS="$OBJS $o"
done
nvcc -shared -o $OUT/libtiresias_gpu_relaxed.so $OBJS -lpthread -ldl
rm -f $OBJS
echo built $OUT/libtiresias_gpu_relaxed.so
