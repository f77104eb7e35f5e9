import os
import subprocess
import sys

import pytest

# The p2p tests run several ranks of the NVLink exchange as contexts of ONE process; on a single-GPU box they share
# the device, and their kernels wait for one another.  More hardware queues than the default 8 keep the ranks'
# streams from being serialised behind each other (must be set before CUDA initialises).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def emul():
    """Host emulation of the extraction kernel's per-thread phases (tests/emul) -- test only."""
    import ctypes as C
    so = os.path.join(ROOT, "tests", "_build", "libtir_emul.so")
    srcs = [os.path.join(ROOT, "tests", "emul", "emul_extract.cpp")] + [
        os.path.join(ROOT, "asterisk_tiresias_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "asterisk_tiresias_b200", "csrc"))
        if f.endswith((".cuh", ".h", ".cpp"))]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["sh", os.path.join(ROOT, "tests", "emul", "build_emul.sh")])
    L = C.CDLL(so)
    L.emul_extract.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
    L.emul_extract_win.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
    L.emul_tables.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.emul_log10f.restype = C.c_float
    L.emul_log10f.argtypes = [C.c_float]
    L.emul_quantize.restype = C.c_int32
    L.emul_quantize.argtypes = [C.c_double]
    L.emul_log10f_sweep.restype = C.c_uint64
    L.emul_log10f_sweep.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    return L


@pytest.fixture(scope="session")
def gpu_ctx():
    from asterisk_tiresias_b200 import capi
    ctx = capi.Context(device=0)
    yield ctx
    ctx.close()
