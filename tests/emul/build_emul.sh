#!/bin/sh
# TEST ONLY: host emulation of the kernel phase functions (see emul_extract.cpp)
set -e
cd "$(dirname "$0")/../.."
mkdir -p tests/_build
g++ -O2 -g -std=c++17 -fPIC -shared -ffp-contract=off -mfma -o tests/_build/libtir_emul.so \
    tests/emul/emul_extract.cpp asterisk_tiresias_b200/csrc/tir_tables.cpp
