"""Quick GPU probe: match parity vs the SQLite oracle on random DBs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from asterisk_tiresias_b200 import capi, synth_db
from oracle import pyoracle as po

ctx = capi.Context(device=0)
rng = np.random.default_rng(5)
for case, (n_audio, null_frac) in enumerate([(50, 0.0), (3000, 0.02), (40000, 0.0)]):
    db = synth_db.make_db(n_audio, 20, 60, seed=case, null_frac=null_frac)
    t = time.time()
    sq = po.SqliteDB()
    for u, y in db:
        sq.add_audio(u, y)
    t_sql = time.time() - t
    ctx.db_load(*synth_db.db_arrays(db))
    print("case", case, "audios", n_audio, "rows", sq.count_rows(), "sqlite load s", round(t_sql, 2), "stats", ctx.db_stats(), flush=True)
    for coefs, tol, lo, hi in [(1, 0.001, -1, -1), (1, 0.01, -1, -1), (2, 0.5, -1, -1), (1, -1.0, 40, 70), (2, 2.0, 30, 60), (1, 0.0, -1, -1)]:
        nq = 12 if n_audio > 10000 else 30
        ok = 0
        for qi in range(nq):
            if qi % 3 == 0:
                y = db[int(rng.integers(0, n_audio))][1].copy()
            else:
                y = synth_db.random_y(rng, int(rng.integers(1, 100)), null_frac=0.05 if qi % 5 == 0 else 0.0)
            exp = sq.search(y, coefs, tol, lo, hi)
            hit = ctx.match(y, None, coefs, tol, lo, hi)[0]
            got = None if hit["match_count"] == 0 else (capi.bytes_to_uuid(hit["uuid"]), int(hit["match_count"]), int(hit["frame_count"]))
            e = None if exp is None else (exp["uuid"], exp["match_count"], exp["frame_count"])
            if got == e:
                ok += 1
            else:
                print("  MISMATCH", coefs, tol, lo, hi, "q", qi, "gpu", got, "sqlite", e)
        print("  coefs", coefs, "tol", tol, "ign", lo, hi, ":", ok, "/", nq, flush=True)
