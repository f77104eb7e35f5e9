"""Build libtiresias_gpu.so (sm_100a only) in-tree with nvcc.  No torch, no JIT cache."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtiresias_gpu.so")
TOOL = os.path.join(HERE, "..", "tools", "tir_concurrent_bench.bin")
HOST_LIB = os.path.join(HERE, "libtiresias_host.so")
SOURCES = ["tir_api.cu", "tir_extract.cu", "tir_match.cu", "tir_p2p.cu", "tir_stream.cu", "tir_tables.cpp", "tir_batcher.cpp", "tir_sqlite.cpp", "tir_group.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 unless told not to; every fused
    # multiply-add of the extraction path is written explicitly (tir_fp.cuh)
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall", "-Xptxas", "-v",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith((".o", ".log"))] + [
        os.path.join(HERE, "..", "include", "tiresias_gpu.h"), os.path.join(HERE, "..", "include", "fp_handler_gpu.h"),
        os.path.join(HERE, "host", "fp_handler_gpu.cpp"), os.path.join(HERE, "..", "tools", "tir_concurrent_bench.cpp")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    objs = []
    log = []
    for src in SOURCES:
        obj = os.path.join(CSRC, os.path.splitext(src)[0] + ".o")
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed on " + src)
        objs.append(obj)
    cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-lpthread", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    # host-side mirror of src/fp_handler.h on top of the C ABI (asterisk_tiresias_b200/host)
    host_src = os.path.join(HERE, "host", "fp_handler_gpu.cpp")
    if os.path.exists(host_src):
        r = subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", host_src, "-o", HOST_LIB, "-L" + HERE,
                            "-ltiresias_gpu", "-lcrypto", "-ldl", "-Wl,-rpath,$ORIGIN",
                            "-Wl,-Bsymbolic"],  # (the drop-in module of the tests exports the same fp_* names)
                           capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("building libtiresias_host.so failed")
        log.append(r.stderr)
    # C++ driver of the concurrent-channel bench (plain C ABI client; tools/tir_concurrent_bench.cpp)
    tool_src = os.path.join(HERE, "..", "tools", "tir_concurrent_bench.cpp")
    if os.path.exists(tool_src):
        r = subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", tool_src, "-o", TOOL, "-L" + HERE, "-ltiresias_gpu",
                            "-Wl,-rpath,$ORIGIN/../asterisk_tiresias_b200"], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("building tools/tir_concurrent_bench failed")
    with open(os.path.join(CSRC, "ptxas.log"), "w") as f:
        # registers / spills / shared memory per kernel; compile times would only make the file churn
        f.write("\n".join(l for l in "\n".join(log).splitlines() if "Compile time" not in l) + "\n")
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
