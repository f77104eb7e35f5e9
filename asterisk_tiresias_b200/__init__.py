"""B200-native fingerprint hot path of asterisk-tiresias (extraction + match) behind a C ABI.

The product is the shared library built from csrc/ (libtiresias_gpu.so, declared in
include/tiresias_gpu.h).  This Python package is only the ctypes binding used by tests/ and
bench.py, plus synthetic input generators; it contains no compute of its own.
"""
