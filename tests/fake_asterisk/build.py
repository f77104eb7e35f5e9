"""Build tests/_build/app_tiresias_dropin.so (TEST INFRASTRUCTURE): the reference's UNCHANGED Asterisk-facing files,
compiled from where they lie under /root/reference/src (nothing is copied into the repo), + the replacement
asterisk_tiresias_b200/host/fp_handler.c + the fake Asterisk runtime, linked against libtiresias_gpu.so and the
image's real libjansson / libsqlite3 / libuuid / libcrypto.  /root/reference does not exist on the GPU box: the
built .so travels with the snapshot (tests/_build is git-ignored, not gpurun-ignored)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
REF = os.environ.get("TIRESIAS_REFERENCE_SRC", "/root/reference/src")
OUT = os.path.join(ROOT, "tests", "_build", "app_tiresias_dropin.so")
UNCHANGED = ["application_handler.c", "cli_handler.c", "app_tiresias.c", "db_ctx_handler.c"]
OURS = [os.path.join(ROOT, "asterisk_tiresias_b200", "host", "fp_handler.c"), os.path.join(HERE, "fake_asterisk.c")]


def reference_present() -> bool:
    return all(os.path.exists(os.path.join(REF, f)) for f in UNCHANGED + ["fp_handler.h", "db_ctx_handler.h", "app_tiresias.h"])


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = OURS + [os.path.join(REF, f) for f in UNCHANGED] + [os.path.join(ROOT, "include", "tiresias_gpu.h")]
    for d, _, fs in os.walk(os.path.join(HERE, "include")):
        deps += [os.path.join(d, f) for f in fs]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False) -> str | None:
    """-> path of the .so, or None when neither the reference sources nor a prebuilt .so exist."""
    if not reference_present():
        return OUT if os.path.exists(OUT) else None
    if not force and not needs_build():
        return OUT
    from asterisk_tiresias_b200 import build as b
    b.build()
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    libdir = os.path.join(ROOT, "asterisk_tiresias_b200")
    cmd = ["gcc", "-shared", "-fPIC", "-O1", "-g", "-std=gnu99", "-pthread", "-Wall", "-Wno-unused-variable", "-Wno-unused-but-set-variable",
           '-DAST_MODULE="app_tiresias"', "-DAST_MODULE_SELF_SYM=__app_tiresias",     # src/Makefile:9
           "-I" + os.path.join(HERE, "include"), "-I" + REF, "-I" + os.path.join(ROOT, "include"),
           *[os.path.join(REF, f) for f in UNCHANGED], *OURS, "-o", OUT,
           "-L" + libdir, "-ltiresias_gpu", "-l:libjansson.so.4", "-l:libsqlite3.so.0", "-l:libuuid.so.1", "-lcrypto", "-lm",
           "-Wl,-rpath,$ORIGIN/../../asterisk_tiresias_b200",
           "-Wl,-Bsymbolic"]     # libtiresias_host.so exports the same fp_* names: no interposition between the two in one process
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("building the drop-in module against the fake Asterisk failed")
    with open(os.path.join(os.path.dirname(OUT), "app_tiresias_dropin.build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
