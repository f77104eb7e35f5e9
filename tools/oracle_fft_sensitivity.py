"""FFT-order sensitivity of the reference's "frame hash" (TEST / ANALYSIS TOOL; uses oracle/).

aubio delegates its FFT to FFTW3f or Ooura depending on the distro build, so two legitimate libaubio builds differ
in float32 rounding.  This script runs the oracle pipeline (identical window, filterbank, log10f, DCT, "%f" ...)
over the bench corpus's kind of clips with three FFTs -- TIR-FFT (what the GPU kernel reproduces), an Ooura-style
radix-2 + rftfsub order, and a float64 FFT rounded to float32 -- and reports, pairwise, the rates SURVEY H2 names:
exact micro-unit hash identity, trunc(max1) identity (what the query side consumes, src/fp_handler.c:290), window
membership at the default tolerance (what the DB side consumes), and the MFCC relative error.

    python tools/oracle_fft_sensitivity.py [n_clips] [seconds]  ->  one JSON object on stdout
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asterisk_tiresias_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

NAMES = {0: "tir_fft", 1: "ooura_style_radix2", 2: "float64_rounded"}


def rates(a, b):
    (ca, va), (cb, vb) = a, b
    ka, kb = np.trunc(va[:, 0] / 1e6), np.trunc(vb[:, 0] / 1e6)
    ina = np.abs(va[:, 0] - np.rint(va[:, 0] / 1e6) * 1e6) <= 1000
    inb = np.abs(vb[:, 0] - np.rint(vb[:, 0] / 1e6) * 1e6) <= 1000
    rel = np.abs(ca.astype(np.float64) - cb) / np.maximum(np.abs(cb), 1e-3 * np.abs(cb).max())
    return {"frames": int(va.shape[0]), "hash_identical": float((va == vb).all(axis=1).mean()), "hash_max1_identical": float((va[:, 0] == vb[:, 0]).mean()),
            "trunc_max1_identical": float((ka == kb).mean()), "trunc_max1_flips": int((ka != kb).sum()),
            "window_membership_identical": float((ina == inb).mean()), "window_membership_flips": int((ina != inb).sum()),
            "mfcc_max_rel_err": float(rel.max()), "mfcc_p999_rel_err": float(np.quantile(rel, 0.999)),
            "max_abs_hash_delta_micro": int(np.abs(va.astype(np.int64) - vb.astype(np.int64)).max())}


def run(n_clips=500, seconds=30.0, sr=8000, n_threads=None):
    pcm, off = synth.make_corpus(n_clips, seconds, sr, ulaw=True, first_index=100003)
    plan = po.Plan(512, 256, 40, 2, sr)
    out = {}
    try:
        for k in NAMES:
            po.set_fft_kind(k)
            c, _, v = plan.extract_batch(pcm, off, n_threads=n_threads or os.cpu_count() or 1, want_y=False)
            out[k] = (c, v)
    finally:
        po.set_fft_kind(0)
    res = {"clips": n_clips, "seconds_per_clip": seconds, "samplerate": sr, "corpus": "asterisk_tiresias_b200.synth.make_corpus (tone/noise/chirp/composite, mu-law round trip)",
           "pairs": {}}
    for a, b in ((0, 1), (0, 2), (1, 2)):
        res["pairs"][f"{NAMES[a]} vs {NAMES[b]}"] = rates(out[a], out[b])
    return res


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    s = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
    print(json.dumps(run(n, s), indent=1))
