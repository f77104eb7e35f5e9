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


def row_major_model_coefs2(db, y, tol):
    """The order of evaluation of tir_match2_kernel (coefs = 2, short queries), restated in numpy: the
    query's windows sorted by (max1 window, lo2, hi2); inside one max1 group both max2 bounds ascend, so the frames a
    stored row matches are a contiguous range found by two binary searches; a matching row is kept as (uuid, first frame,
    last frame); the union of a uuid's ranges is its vote count (one vote per frame and uuid: GROUP BY audio_uuid)."""
    q = lambda v: int(synth_db.quantize_y(np.array([v]))[0])
    wins = []
    for v1, v2 in np.where(np.isfinite(y), y, 0.0):
        f = float(np.trunc(v1))
        wins.append((q(f - tol), q(f + tol), max(q(v2 - tol), -(2**31) + 1), q(v2 + tol)))
    wins.sort()
    groups = {}                                                        # (lo1, hi1) -> positions in the sorted query
    for pos, w in enumerate(wins):
        groups.setdefault(w[:2], []).append(pos)
    best = None
    for u, ydb in db:
        v = synth_db.quantize_y(ydb)
        frames = set()
        for (lo1, hi1), poss in groups.items():
            lo2 = np.array([wins[p][2] for p in poss]); hi2 = np.array([wins[p][3] for p in poss])
            assert (np.diff(lo2) >= 0).all() and (np.diff(hi2) >= 0).all()      # what the binary searches rely on
            for r in np.nonzero((v[:, 0] != -(2**31)) & (v[:, 0] >= lo1) & (v[:, 0] <= hi1))[0]:
                a = int(np.searchsorted(hi2, v[r, 1], side="left"))             # first frame with hi2 >= max2
                b = int(np.searchsorted(lo2, v[r, 1], side="right"))            # first frame with lo2 > max2
                frames.update(poss[a:b])
        if frames and (best is None or (len(frames), u) > best):
            best = (len(frames), u)
    return None if best is None else (best[1], best[0], y.shape[0])


def test_row_major_coefs2_evaluation_order_equals_sql(oracle):
    rng = np.random.default_rng(8)
    db = synth_db.make_db(250, 3, 25, seed=21, near_int_frac=0.7, null_frac=0.04)
    db.append((synth_db.uuid_for(9_100_001), np.array([[17.0003, 5.0], [17.0004, 5.0002], [16.9998, 5.0001], [18.0002, 7.5]])))  # rows sharing frames
    sq = oracle.SqliteDB()
    for u, y in db:
        sq.add_audio(u, y)
    queries = [synth_db.random_y(rng, int(rng.integers(1, 60)), near_int_frac=0.7) for _ in range(10)]
    queries += [db[3][1].copy(), np.array([[17.4, 5.0001], [17.9, 5.0], [17.2, 4.99995], [18.1, 7.5002], [16.0, 5.0]])]
    for tol in (0.001, 0.02, 0.6):
        for y in queries:
            h = sq.search(y, 2, tol, has_y=np.isfinite(y))
            want = None if h is None else (h["uuid"], h["match_count"], h["frame_count"])
            assert row_major_model_coefs2(db, y, tol) == want, tol


def _pat_hash(p):
    """tir_pat_hash (csrc/tir_match.cu), restated"""
    M = (1 << 64) - 1
    p ^= p >> 33
    p = (p * 0xff51afd7ed558ccd) & M
    p ^= p >> 33
    return ((p & 0xFFFFFFFF) ^ ((p >> 17) & 0xFFFFFFFF)) & 0xFFFFFFFF


def test_pattern_hash_spreads_sparse_bit_sets():
    """The patterns are one-, two- and three-bit sets over up to 64 windows.  The first hash (a 32-bit multiply, then the
    LOW product bits) sent every pattern made of windows >= 17 to slot 0 of the 1 024-slot CTA table -- K = 32 took 219 us
    per batch instead of 84.  The mixer must spread them: at most a handful of patterns per slot, and the old one shown
    failing the same bound."""
    pats = [1 << a for a in range(64)] + [(1 << a) | (1 << b) for a in range(64) for b in range(a)]       # 2 080 patterns
    load = np.bincount([_pat_hash(p) & 1023 for p in pats], minlength=1024)
    assert load.max() <= 10 and (load > 0).sum() >= 850, (load.max(), (load > 0).sum())
    old = np.bincount([(((p & 0xFFFFFFFF) * 0x9e3779b1) & 0xFFFFFFFF) >> 7 & 1023 for p in pats if p < (1 << 32)], minlength=1024)
    assert old.max() > 100                                               # the cluster the round-1 hash built at slot 0
    g = np.bincount([_pat_hash(p) & 16383 for p in pats], minlength=16384)
    assert g.max() <= 4


def _group_bound(key, lo, hi, target, upper, G):
    """tir_group_bound (csrc/tir_match.cu), one group of G lanes, restated: (G+1)-ary search, then one step of G probes"""
    levels = 0
    while hi - lo > G:
        step = (hi - lo + G) // (G + 1)
        nb = 0
        for sub in range(G):                                             # probes are monotone: the first nb are below
            p = lo + (sub + 1) * step - 1
            if p < hi and (key[p] <= target if upper else key[p] < target):
                nb += 1
        lo, hi = lo + nb * step, (min(hi, lo + (nb + 1) * step) if nb < G else hi)
        levels += 1
    nb = sum(1 for sub in range(G) if lo + sub < hi and (key[lo + sub] <= target if upper else key[lo + sub] < target))
    return lo + nb, levels


def test_lane_group_bound_search_equals_searchsorted():
    rng = np.random.default_rng(12)
    for n in (0, 1, 2, 31, 32, 33, 1000, 40_000):
        key = np.sort(rng.integers(-50, 50, n) * 1000 + rng.integers(0, 3, n))       # many duplicates
        for G in (2, 4, 8, 16, 32):
            for t in list(rng.integers(-60_000, 60_000, 12)) + ([int(key[0]), int(key[-1]), int(key[n // 2])] if n else []):
                lo, hi = (5, n - 3) if n > 20 else (0, n)                             # a block is a sub-range of the array
                got_l, lv = _group_bound(key, lo, hi, t, False, G)
                got_u, _ = _group_bound(key, lo, hi, t, True, G)
                assert got_l == lo + int(np.searchsorted(key[lo:hi], t, side="left")), (n, G, t)
                assert got_u == lo + int(np.searchsorted(key[lo:hi], t, side="right")), (n, G, t)
                assert lv <= int(np.ceil(np.log(max(hi - lo, 2)) / np.log(G + 1))) + 1
