"""The drop-in boundary without a GPU: the reference's UNCHANGED application_handler.c, cli_handler.c,
app_tiresias.c and db_ctx_handler.c compile (from /root/reference/src, in place) and link against the
replacement asterisk_tiresias_b200/host/fp_handler.c, which exports the 13 functions of src/fp_handler.h:13-38;
and with no CUDA device the module DECLINES to load (src/app_tiresias.c:585-589) -- there is no CPU fallback."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from fake_asterisk import build as fa_build  # noqa: E402

FP_API = ["fp_init", "fp_term", "fp_create_context_list_info", "fp_delete_context_list_info", "fp_get_context_lists_all",
          "fp_get_context_list_info", "fp_get_audio_lists_all", "fp_get_audio_lists_by_contextname", "fp_craete_audio_list_info",
          "fp_delete_audio_list_info", "fp_search_fingerprint_info", "fp_generate_uuid", "fp_create_hash"]


@pytest.fixture(scope="module")
def dropin():
    so = fa_build.build()
    if so is None:
        pytest.skip("neither /root/reference/src nor a prebuilt tests/_build/app_tiresias_dropin.so")
    return so


def test_unchanged_reference_files_link_against_the_replacement(dropin):
    L = C.CDLL(dropin)
    for name in FP_API + ["fake_ast_module_info", "cli_init", "cli_term", "application_init", "application_term", "db_ctx_init"]:
        assert hasattr(L, name), name
    if fa_build.reference_present():
        # the prototypes of the replacement are the reference header's: fp_handler.c includes that very header
        # (a mismatch is a compile error) and declares nothing else public beside g_db_ctx
        hdr = open(os.path.join(fa_build.REF, "fp_handler.h")).read()
        assert sorted(re.findall(r"\b(fp_[a-z_]+)\s*\(", hdr)) == sorted(FP_API)
        log = open(os.path.join(ROOT, "tests", "_build", "app_tiresias_dropin.build.log")).read()
        for f in fa_build.UNCHANGED:
            assert os.path.join(fa_build.REF, f) in log.splitlines()[0]          # compiled from where they lie
        out = subprocess.run(["nm", "-D", "--defined-only", dropin], capture_output=True, text=True).stdout
        exported_fp = sorted(set(re.findall(r" T (fp_[a-z_]+)$", out, re.M)))
        assert exported_fp == sorted(FP_API)


def test_module_declines_to_load_without_a_gpu(dropin, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible: covered by tests/test_gpu_asterisk_dropin.py")
    code = f"""
import sys
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {ROOT!r} + '/tests')
from fake_asterisk.harness import FakeAsterisk
fa = FakeAsterisk({str(tmp_path)!r})
fa.write_conf("[global]\\ntolerance=0.01\\n[ctx]\\ndirectory={tmp_path}/audio\\n")
rc = fa.load()
print("RC", rc)
print(fa.log())
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "RC 1" in r.stdout, r.stdout + r.stderr        # AST_MODULE_LOAD_DECLINE
    assert "Could not initiate the GPU fingerprint engine" in r.stdout
    assert "Could not initiate fp_handler" in r.stdout    # the module shell's own message (src/app_tiresias.c:92-96)
