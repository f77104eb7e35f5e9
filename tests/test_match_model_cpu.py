"""CPU check of the ALGORITHM behind the shared-window match path, against the
reference's SQL on the real SQLite: the batch's distinct windows are scanned once, every audio gets
a bit pattern (bit k: it has a row in window k -- GROUP BY audio_uuid makes it a bit, not a count),
only the greatest uuid per pattern can win, and a query weighs the patterns with the multiplicities
of its windows:  match_count(q, uuid) = sum_k weight(q, k) * bit_k(pattern(uuid)).
A numpy restatement of that order of evaluation (test code: the kernels are checked on the GPU)."""
import numpy as np

from asterisk_tiresias_b200 import synth_db


def shared_window_model(db, queries, tol):
    """db: list of (uuid text, y[F,2]); queries: list of y[F,2]  ->  list of (uuid, match_count, frame_count) | None"""
    q_win = []                                                        # per query: {(lo, hi): weight}
    for y in queries:
        w = {}
        for v in np.where(np.isfinite(y[:, 0]), y[:, 0], 0.0):
            f = float(np.trunc(v))
            key = (int(synth_db.quantize_y(np.array([f - tol]))[0]), int(synth_db.quantize_y(np.array([f + tol]))[0]))
            w[key] = w.get(key, 0) + 1
        q_win.append(w)
    distinct = sorted({k for w in q_win for k in w})                  # the batch's window set
    bit = {k: i for i, k in enumerate(distinct)}
    greatest = {}                                                     # pattern -> greatest uuid carrying it
    for u, y in db:
        v1 = synth_db.quantize_y(y[:, 0])
        ok = v1 != -(2**31)
        p = 0
        for k, (lo, hi) in enumerate(distinct):
            if np.any(ok & (v1 >= lo) & (v1 <= hi)):
                p |= 1 << k
        if p and (p not in greatest or u > greatest[p]):
            greatest[p] = u
    out = []
    for y, w in zip(queries, q_win):
        best = None
        for p, u in greatest.items():
            score = sum(wt for k, wt in w.items() if (p >> bit[k]) & 1)
            if score and (best is None or (score, u) > best):
                best = (score, u)
        out.append(None if best is None else (best[1], best[0], y.shape[0]))
    return out


def test_shared_window_evaluation_order_equals_sql(oracle):
    rng = np.random.default_rng(3)
    db = synth_db.make_db(400, 3, 25, seed=13, lo=10.0, hi=24.0, near_int_frac=0.7, null_frac=0.03)
    db += [(synth_db.uuid_for(9_000_000 + i), db[i][1].copy()) for i in range(12)]     # ties -> greatest uuid
    sq = oracle.SqliteDB()
    for u, y in db:
        sq.add_audio(u, y)
    queries = [synth_db.random_y(rng, int(rng.integers(1, 40)), lo=10.0, hi=24.0, near_int_frac=0.7, null_frac=0.05)
               for _ in range(25)] + [db[7][1].copy(), db[405][1].copy()]
    for tol in (0.001, 0.03):
        got = shared_window_model(db, queries, tol)
        n_windows = len({float(np.trunc(v)) for y in queries for v in np.where(np.isfinite(y[:, 0]), y[:, 0], 0.0)})
        assert n_windows > 12                        # the batch needs the hashed pattern tables on the GPU
        for y, g in zip(queries, got):
            h = sq.search(y, 1, tol, has_y=np.isfinite(y))
            want = None if h is None else (h["uuid"], h["match_count"], h["frame_count"])
            assert g == want
