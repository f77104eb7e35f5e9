"""Oracle (and the host-run kernel phases) against the committed golden fixtures."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_reproduces_extract_golden(oracle, emul):
    g = np.load(os.path.join(GOLD, "extract_golden.npz"))
    p = oracle.Plan()
    coef, y, vq = p.extract_batch(g["pcm"], g["clip_off"], n_threads=2)
    assert np.array_equal(coef.view(np.uint32), g["coef_bits"])
    assert np.array_equal(vq, g["vq"])
    off = g["clip_off"].astype(np.int64)
    fo = np.concatenate([[0], np.cumsum((np.diff(off) + 255) // 256)])
    for c in range(off.size - 1):
        pcm = np.ascontiguousarray(g["pcm"][off[c]:off[c + 1]])
        F = int(fo[c + 1] - fo[c])
        c2 = np.zeros((F, 2), np.float32); v2 = np.zeros((F, 2), np.int32)
        assert emul.emul_extract(pcm.ctypes.data, pcm.size, 8000, c2.ctypes.data, v2.ctypes.data) == 0
        assert np.array_equal(c2.view(np.uint32), g["coef_bits"][fo[c]:fo[c + 1]])
        assert np.array_equal(v2, g["vq"][fo[c]:fo[c + 1]])


def test_sqlite_reproduces_match_golden(oracle):
    g = np.load(os.path.join(GOLD, "match_golden.npz"))
    db = oracle.SqliteDB()
    off = g["db_off"].astype(np.int64)
    for i, u in enumerate(g["db_uuid"]):
        db.add_audio(str(u), g["db_y"][off[i]:off[i + 1]])
    qo = g["q_off"].astype(np.int64)
    for i in range(qo.size - 1):
        coefs, tol, lo, hi = g["params"][g["exp_param"][i]]
        y = g["q_y"][qo[i]:qo[i + 1]]
        hit = db.search(y, int(coefs), float(tol), int(lo), int(hi), has_y=np.isfinite(y))
        if g["exp_count"][i] == 0:
            assert hit is None
        else:
            assert (hit["uuid"], hit["match_count"], hit["frame_count"]) == (str(g["exp_uuid"][i]), g["exp_count"][i], g["exp_frames"][i])
