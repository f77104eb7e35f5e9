"""Regenerate the golden fixtures from the CPU oracle (run here, in the build container):
    python tests/golden/make_golden.py
extract_golden.npz : PCM16 clips + oracle coefficients (bit patterns) + micro-unit values
match_golden.npz   : a small fingerprint DB, queries, parameters and the (uuid, match_count,
                     frame_count) the real SQLite returned for the reference's SQL text
The reference ships no fixtures of its own; these pin OUR oracle so that a
later change to oracle/ or to the kernels cannot drift silently."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from asterisk_tiresias_b200 import synth, synth_db  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def extract_golden():
    plan = po.Plan()
    kinds = ["tone", "noise", "chirp", "composite", "silence", None, None, None, None, None]
    secs = [0.5, 0.5, 0.5, 0.41, 0.1, 0.003, 0.033, 0.7, 0.25, 1.1]
    clips = [synth.make_clip(100 + i, secs[i], kind=kinds[i], ulaw=(i % 3 == 1)) for i in range(len(kinds))]
    clips.append(np.array([-32768, 32767, 0, 1, -1] * 101, np.int16))       # extremes, odd length
    clips.append(np.zeros(0, np.int16))                                       # empty clip
    off = np.zeros(len(clips) + 1, np.uint64)
    off[1:] = np.cumsum([c.size for c in clips])
    pcm = np.concatenate(clips)
    coef, y, vq = plan.extract_batch(pcm, off)
    np.savez_compressed(os.path.join(HERE, "extract_golden.npz"), pcm=pcm, clip_off=off, coef_bits=coef.view(np.uint32), vq=vq)
    print("extract golden:", len(clips), "clips", coef.shape[0], "frames")


def match_golden():
    rng = np.random.default_rng(77)
    db = synth_db.make_db(300, 10, 40, seed=7, null_frac=0.02)
    # make ties likely: duplicate the rows of some audios under other uuids
    for i in range(20):
        db.append((synth.uuid_for(900000 + i), db[i][1].copy()))
    sq = po.SqliteDB()
    for u, y in db:
        sq.add_audio(u, y)
    params = [(1, 0.001, -1, -1), (1, 0.01, -1, -1), (2, 0.5, -1, -1), (1, -1.0, 40, 70), (2, 2.0, 30, 60), (1, 0.0, -1, -1),
              (2, 0.001, -1, -1), (1, 0.3, 60, -1)]
    queries, expect = [], []
    for pi, (coefs, tol, lo, hi) in enumerate(params):
        for qi in range(10):
            if qi % 3 == 0:
                y = db[int(rng.integers(0, len(db)))][1].copy()
            else:
                y = synth_db.random_y(rng, int(rng.integers(1, 80)), null_frac=0.1 if qi % 4 == 1 else 0.0)
            hit = sq.search(y, coefs, tol, lo, hi, has_y=np.isfinite(y))
            queries.append(y)
            expect.append((pi, "" if hit is None else hit["uuid"], 0 if hit is None else hit["match_count"], y.shape[0]))
    foff = np.zeros(len(queries) + 1, np.uint64)
    foff[1:] = np.cumsum([q.shape[0] for q in queries])
    np.savez_compressed(
        os.path.join(HERE, "match_golden.npz"),
        db_uuid=np.array([u for u, _ in db]), db_off=np.cumsum([0] + [y.shape[0] for _, y in db]).astype(np.uint64),
        db_y=np.concatenate([y for _, y in db]), params=np.array(params, np.float64), q_y=np.concatenate(queries), q_off=foff,
        exp_param=np.array([e[0] for e in expect], np.int32), exp_uuid=np.array([e[1] for e in expect]),
        exp_count=np.array([e[2] for e in expect], np.int32), exp_frames=np.array([e[3] for e in expect], np.int32),
        sqlite_version=np.array(po.SqliteDB.sqlite_version()))
    print("match golden:", len(db), "audios", len(queries), "queries", sum(1 for e in expect if e[1]), "found")


if __name__ == "__main__":
    extract_golden()
    match_golden()
