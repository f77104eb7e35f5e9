"""Synthetic fingerprint DBs / queries for the match tests and benches (numpy; not product code).

Values are the y = 10*log10|c| doubles the reference stores.  With the reference's default
tolerance (0.001) a stored frame only matches when its max1 lies within 0.001 of an integer, so
the generators put a controllable share of the values next to integers.
"""
from __future__ import annotations

import numpy as np

from .synth import SEED0, uuid_for


def random_y(rng, n_frames, near_int_frac=0.3, lo=15.0, hi=19.0, spread=0.0015, null_frac=0.0):
    y1 = rng.uniform(lo, hi, n_frames)
    near = rng.random(n_frames) < near_int_frac
    y1[near] = np.round(y1[near]) + rng.uniform(-spread, spread, int(near.sum()))
    y2 = rng.uniform(-5.0, 20.0, n_frames)
    y = np.stack([y1, y2], axis=1)
    if null_frac > 0:
        m = rng.random((n_frames, 2)) < null_frac
        y[m] = np.nan  # non-finite -> column NULL
    return y


def make_db(n_audio, frames_lo=20, frames_hi=60, seed=0, **kw):
    """-> list of (uuid_text, y[F,2]) ."""
    rng = np.random.default_rng(SEED0 + 1000 + seed)
    out = []
    for a in range(n_audio):
        F = int(rng.integers(frames_lo, frames_hi + 1))
        out.append((uuid_for(seed * 1_000_003 + a), random_y(rng, F, **kw)))
    return out


def quantize_y(y):
    """float64 y -> int32 micro-units as "%f" would print them (NaN/inf -> NULL).  Host helper for
    loading test DBs; exact for the test ranges (checked against the oracle in tests)."""
    from decimal import Decimal, ROUND_HALF_EVEN
    y = np.asarray(y, dtype=np.float64)
    flat = y.reshape(-1)
    out = np.empty(flat.size, np.int64)
    for i, v in enumerate(flat):
        if not np.isfinite(v):
            out[i] = -(2**31)
        else:
            out[i] = int((Decimal(float(v)) * 1000000).quantize(Decimal(1), rounding=ROUND_HALF_EVEN))
    return out.reshape(y.shape).astype(np.int32)


def db_arrays(db):
    """list of (uuid, y) -> (uuid bytes [n,16], row_off [n+1], v1, v2) for tir_db_load."""
    import uuid as _uuid
    uu = np.stack([np.frombuffer(_uuid.UUID(u).bytes, np.uint8) for u, _ in db]) if db else np.zeros((0, 16), np.uint8)
    row_off = np.zeros(len(db) + 1, np.uint64)
    row_off[1:] = np.cumsum([y.shape[0] for _, y in db])
    if db:
        yy = np.concatenate([y for _, y in db])
        v = quantize_y(yy)
    else:
        v = np.zeros((0, 2), np.int32)
    return uu, row_off, np.ascontiguousarray(v[:, 0]), np.ascontiguousarray(v[:, 1])
