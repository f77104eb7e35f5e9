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


def test_host_mirror_exports_the_reference_interface_and_refuses_to_run_without_a_gpu():
    """libtiresias_host.so carries every function of src/fp_handler.h:13-38 under the same name; with no
    CUDA device fp_init() must fail (the module would decline to load) instead of falling back."""
    import torch
    from asterisk_tiresias_b200.host import fp_host
    L = fp_host.lib()
    for name in ["fp_init", "fp_term", "fp_create_context_list_info", "fp_delete_context_list_info", "fp_get_context_lists_all",
                 "fp_get_context_list_info", "fp_get_audio_lists_all", "fp_get_audio_lists_by_contextname",
                 "fp_craete_audio_list_info", "fp_delete_audio_list_info", "fp_search_fingerprint_info", "fp_generate_uuid",
                 "fp_create_hash"]:
        assert hasattr(L, name), name
    u = fp_host.fp_generate_uuid()
    assert len(u) == 36 and u[14] == "4"
    if not torch.cuda.is_available():
        assert fp_host.fp_init(None, 0) is False
        assert fp_host.fp_init(None, 0) is False          # and a failed load leaves no half-open state behind
        assert fp_host.fp_get_audio_lists_all() == []
        assert fp_host.fp_search_fingerprint_info("ctx", "/nonexistent.wav") is None


def test_batcher_and_stream_entry_points_reject_bad_arguments_without_a_gpu():
    from asterisk_tiresias_b200 import capi
    L = capi.lib()
    assert L.tir_batcher_start(None, 16, 100) != 0
    assert L.tir_search_one(None, None, 0, 1, 0.001, -1, -1, None) != 0
    assert L.tir_stream_feed(None, None, 0) != 0
    assert L.tir_db_load_sqlite(None, None, None, None, None) != 0
