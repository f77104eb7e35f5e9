"""The SQLite seams of the product (tir_sqlite.cpp): batched ingest must store what the reference's
per-frame textual INSERTs store, and the device table loaded from a SQLite connection must answer
like SQLite itself."""
import numpy as np
import pytest

from asterisk_tiresias_b200 import capi, synth, synth_db


def test_batched_ingest_stores_what_the_text_inserts_store(oracle):
    """src/fp_handler.c:538-575 + src/db_ctx_handler.c:480 (text path, run literally by the oracle)
    vs tir_sqlite_insert_fingerprints (one transaction, bound doubles).  No GPU involved."""
    rng = np.random.default_rng(3)
    plan = oracle.Plan()
    a, b = oracle.SqliteDB(), oracle.SqliteDB()
    cases = [plan.extract(synth.make_clip(i, 2.0))[1] for i in range(6)]             # real y = 10*log10|c|
    cases.append(synth_db.random_y(rng, 300, null_frac=0.1))                          # NULL columns
    cases.append(np.array([[17.0, -0.0], [1e-7, -1e-7], [5e-7, 1.5e-6], [2.5e-6, 16.9999995], [123.456789499, -45.0000005],
                           [0.1 + 0.2, 1 / 3], [np.inf, np.nan], [-np.inf, 2.0]]))
    for i, y in enumerate(cases):
        u = synth.uuid_for(600 + i)
        a.add_audio(u, y, context="ctx-a", literal_autocommit=(i % 2 == 0))
        vq = np.array([[oracle.quantize(v) if np.isfinite(v) else oracle.NULL_V for v in row] for row in y], np.int32)
        capi.sqlite_insert_fingerprints(b.handle, "ctx-a", u, vq)
        fa, fb = a.dump_audio(u), b.dump_audio(u)
        assert np.array_equal(fa[0], fb[0]) and np.array_equal(fa[0], np.arange(y.shape[0]))   # frame_idx
        assert np.array_equal(fa[3], fb[3]) and np.array_equal(fa[4], fb[4])                    # storage classes (REAL / NULL)
        assert np.array_equal(fa[1].view(np.uint64), fb[1].view(np.uint64))                    # max1 bit for bit
        assert np.array_equal(fa[2].view(np.uint64), fb[2].view(np.uint64))
        assert fa[5] == fb[5] == "ctx-a"
    assert a.count_rows() == b.count_rows()
    # and the two databases answer the reference's search SQL identically
    for y in cases[:3]:
        assert a.search(y, 1, 0.01) == b.search(y, 1, 0.01)


@pytest.mark.gpu
def test_device_table_loaded_from_sqlite_answers_like_sqlite(gpu_ctx, oracle, tmp_path):
    rng = np.random.default_rng(8)
    db = synth_db.make_db(1500, 5, 40, seed=77, null_frac=0.02)
    sq = oracle.SqliteDB()
    for u, y in db:
        sq.add_audio(u, y)
    n_audio, n_rows, skipped = gpu_ctx.db_load_sqlite(sq.handle)
    assert (n_audio, n_rows, skipped) == (len(db), sq.count_rows(), 0)
    assert gpu_ctx.db_stats() == (len(db), sq.count_rows())
    for q in range(12):
        y = db[int(rng.integers(0, len(db)))][1] if q % 2 else synth_db.random_y(rng, 60)
        for coefs, tol in ((1, 0.001), (1, 0.02), (2, 0.7)):
            h = gpu_ctx.match(y, None, coefs, tol)[0]
            exp = sq.search(y, coefs, tol, has_y=np.isfinite(y))
            got = None if h["match_count"] == 0 else (capi.bytes_to_uuid(h["uuid"]), int(h["match_count"]), int(h["frame_count"]))
            assert got == (None if exp is None else (exp["uuid"], exp["match_count"], exp["frame_count"]))
    # a row whose audio_uuid is not a uuid (reference quirk K2) is skipped, not fatal
    vq = np.array([[17000000, 0]], np.int32)
    capi.sqlite_insert_fingerprints(sq.handle, "ctx", "/tmp/some-file.wav", vq)
    assert gpu_ctx.db_load_sqlite(sq.handle)[2] == 1
