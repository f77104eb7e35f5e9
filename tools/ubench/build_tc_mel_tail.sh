#!/bin/sh
# builds the tensor-core mel-tail experiment (tools/ubench/tc_mel_tail.cu); run from the repo root
set -e
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -I asterisk_tiresias_b200/csrc \
     -Xptxas -v -o tools/ubench/tc_mel_tail.bin tools/ubench/tc_mel_tail.cu asterisk_tiresias_b200/csrc/tir_tables.cpp
