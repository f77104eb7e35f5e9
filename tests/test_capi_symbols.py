"""The C-ABI library loads and exports every symbol include/tiresias_gpu.h declares (no compute)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tiresias_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tir_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_documented_surface():
    syms = declared_symbols()
    for s in ["tir_open", "tir_close", "tir_extract", "tir_extract_dev", "tir_db_load", "tir_db_add", "tir_db_remove",
              "tir_match", "tir_match_dev", "tir_search", "tir_merge_hits_dev", "tir_shard_of", "tir_last_error"]:
        assert s in syms


def test_library_exports_every_declared_symbol():
    from asterisk_tiresias_b200 import build, capi
    if not os.path.exists(capi.LIB_PATH):
        build.build()
    lib = C.CDLL(capi.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    lib.tir_abi_version.restype = C.c_int
    assert lib.tir_abi_version() == 1
    lib.tir_n_frames.restype = C.c_uint64
    lib.tir_n_frames.argtypes = [C.c_uint64, C.c_int]
    assert lib.tir_n_frames(240000, 256) == 938 and lib.tir_n_frames(0, 256) == 0 and lib.tir_n_frames(1, 256) == 1


def test_shard_of_is_stable_and_balanced():
    import numpy as np
    from asterisk_tiresias_b200 import capi, synth
    if not os.path.exists(capi.LIB_PATH):
        pytest.skip("library not built")
    counts = np.zeros(8, int)
    for i in range(4000):
        u = capi.uuid_to_bytes(synth.uuid_for(i))
        s = capi.shard_of(u, 8)
        assert s == capi.shard_of(u, 8) and 0 <= s < 8
        counts[s] += 1
    assert counts.min() > 400 and capi.shard_of(u, 1) == 0


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    from asterisk_tiresias_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.TirError):
        capi.Context(device=0)
