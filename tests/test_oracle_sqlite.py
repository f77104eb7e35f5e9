"""The match oracle IS SQLite running the reference's SQL text; these tests pin the properties of
that engine the CUDA path relies on (probes)."""
import numpy as np

from asterisk_tiresias_b200 import synth, synth_db


def test_sqlite_version(oracle):
    assert oracle.SqliteDB.sqlite_version().startswith("3.")


def test_inclusive_bounds_and_micro_unit_equivalence(oracle):
    db = oracle.SqliteDB()
    ys = [16.999, 17.001, 17.0010004, 17.0010006, 16.9989994, 16.9989996]
    for i, v in enumerate(ys):
        db.add_audio(synth.uuid_for(i), np.array([[v, 0.0]]))
    # stored through "%f": 16.999000 17.001000 17.001000 17.001001 16.998999 16.999000
    hit = db.search(np.array([[17.3, 0.0]]), coefs=1, tolerance=0.001)
    assert hit["match_count"] == 1 and hit["frame_count"] == 1
    assert hit["votes_total"] == 4   # four audios inside [16.999000, 17.001000]
    v = synth_db.quantize_y(np.array(ys))
    assert int(((v >= 16999000) & (v <= 17001000)).sum()) == 4


def test_tie_goes_to_greatest_uuid(oracle):
    for n in (5, 50, 500):
        db = oracle.SqliteDB()
        uu = [synth.uuid_for(1000 + i) for i in range(n)]
        for u in uu:
            db.add_audio(u, np.array([[17.0, 1.0], [18.0, 1.0]]))
        hit = db.search(np.array([[17.2, 0.0], [18.9, 0.0], [17.9, 0.0]]))
        assert hit["match_count"] == 3 and hit["uuid"] == max(uu)


def test_one_vote_per_frame_and_frame_count(oracle):
    db = oracle.SqliteDB()
    a, b = synth.uuid_for(1), synth.uuid_for(2)
    db.add_audio(a, np.array([[17.0, 0.0]] * 10))          # many rows in the window: still one vote
    db.add_audio(b, np.array([[17.0, 0.0], [18.0, 0.0]]))
    hit = db.search(np.array([[17.5, 0], [17.6, 0], [18.1, 0], [3.0, 0]]))
    assert hit == {"uuid": b, "match_count": 3, "frame_count": 4, "votes_total": 5}


def test_null_and_missing_values(oracle):
    db = oracle.SqliteDB()
    a, b = synth.uuid_for(3), synth.uuid_for(4)
    db.add_audio(a, np.array([[np.nan, 5.0], [17.0, np.nan]]))
    db.add_audio(b, np.array([[0.0, 0.0]]))
    # missing query max1 reads as 0.0 -> matches b's 0.0 row
    hit = db.search(np.array([[np.nan, 0.0]]), has_y=np.array([[0, 1]]))
    assert hit["uuid"] == b
    # coefs=2: NULL max2 never satisfies the max2 predicate
    assert db.search(np.array([[17.2, 5.0]]), coefs=2, tolerance=0.5) is None
    assert db.search(np.array([[17.2, 5.0]]), coefs=1, tolerance=0.5)["uuid"] == a


def test_argument_rules(oracle):
    db = oracle.SqliteDB()
    db.add_audio(synth.uuid_for(9), np.array([[17.0, 2.0]]))
    assert db.search(np.array([[17.0, 2.0]]), coefs=3) is None and db.search(np.array([[17.0, 2.0]]), coefs=0) is None
    assert db.search(np.array([[17.0004, 2.0]]), tolerance=-5)["match_count"] == 1   # < 0 -> 0.001
    # freq_ignore: 10*log10(50) = 16.99 -> frame with trunc 17 kept by low=50, dropped by high=40
    assert db.search(np.array([[17.4, 2.0]]), freq_ignore_low=50) is not None
    assert db.search(np.array([[17.4, 2.0]]), freq_ignore_high=40) is None
    # delete
    db.delete_audio(synth.uuid_for(9))
    assert db.search(np.array([[17.0, 2.0]])) is None and db.count_rows() == 0
