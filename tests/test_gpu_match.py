"""GPU parity tests of the match path through the C ABI against the real SQLite running the
reference's SQL text: identical winner uuid, match_count and frame_count on every query
(bit-exact bar for integer work), ties and NULLs included."""
import os

import numpy as np
import pytest

from asterisk_tiresias_b200 import capi, synth, synth_db

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gpu_result(hit):
    if hit["match_count"] == 0:
        return None
    return (capi.bytes_to_uuid(hit["uuid"]), int(hit["match_count"]), int(hit["frame_count"]))


def sql_result(hit):
    return None if hit is None else (hit["uuid"], hit["match_count"], hit["frame_count"])


def test_golden_fixture(gpu_ctx):
    g = np.load(os.path.join(GOLD, "match_golden.npz"))
    off = g["db_off"]
    v = synth_db.quantize_y(g["db_y"])
    uu = np.stack([capi.uuid_to_bytes(str(u)) for u in g["db_uuid"]])
    gpu_ctx.db_load(uu, off, v[:, 0], v[:, 1])
    qo = g["q_off"].astype(np.int64)
    for pi, (coefs, tol, lo, hi) in enumerate(g["params"]):
        idx = np.nonzero(g["exp_param"] == pi)[0]
        ys = [g["q_y"][qo[i]:qo[i + 1]] for i in idx]
        foff = np.zeros(len(ys) + 1, np.uint64); foff[1:] = np.cumsum([y.shape[0] for y in ys])
        hits = gpu_ctx.match(np.concatenate(ys), foff, int(coefs), float(tol), int(lo), int(hi))   # one batched call
        for h, i in zip(hits, idx):
            exp = None if g["exp_count"][i] == 0 else (str(g["exp_uuid"][i]), int(g["exp_count"][i]), int(g["exp_frames"][i]))
            assert gpu_result(h) == exp, (pi, i)


PARAMS = [(1, 0.001, -1, -1), (1, 0.01, -1, -1), (2, 0.5, -1, -1), (1, -1.0, 40, 70), (2, 2.0, 30, 60), (1, 0.0, -1, -1),
          (2, 0.002, 50, -1), (1, 1e-6, -1, -1), (1, 1.5e-6, -1, -1)]


@pytest.mark.parametrize("n_audio,null_frac,seed", [(40, 0.0, 1), (2500, 0.03, 2), (20000, 0.0, 3)])
def test_differential_vs_sqlite(gpu_ctx, oracle, n_audio, null_frac, seed):
    rng = np.random.default_rng(seed)
    db = synth_db.make_db(n_audio, 5, 30 if n_audio > 10000 else 50, seed=seed, null_frac=null_frac)
    sq = oracle.SqliteDB()
    for u, y in db:
        sq.add_audio(u, y)
    gpu_ctx.db_load(*synth_db.db_arrays(db))
    assert gpu_ctx.db_stats() == (n_audio, sq.count_rows())
    for coefs, tol, lo, hi in PARAMS:
        if n_audio > 10000 and tol > 0.1:
            continue          # keeps the SQLite side of the test in seconds
        ys = []
        for qi in range(8):
            if qi % 3 == 0:
                y = db[int(rng.integers(0, n_audio))][1].copy()
                if qi % 6 == 0:
                    y = y + rng.normal(0, 2e-4, y.shape)     # noisy copy
            else:
                y = synth_db.random_y(rng, int(rng.integers(1, 120)), null_frac=0.1 if qi % 4 == 1 else 0.0)
            ys.append(y)
        foff = np.zeros(len(ys) + 1, np.uint64); foff[1:] = np.cumsum([y.shape[0] for y in ys])
        hits = gpu_ctx.match(np.concatenate(ys), foff, coefs, tol, lo, hi)
        for y, h in zip(ys, hits):
            exp = sq.search(y, coefs, tol, lo, hi, has_y=np.isfinite(y))
            assert gpu_result(h) == sql_result(exp), (coefs, tol, lo, hi)


def test_shared_window_and_per_query_paths_agree(gpu_ctx, oracle):
    """The engine has two evaluation orders for the same votes: a batch with few distinct windows is
    scanned once for all its queries (shared-window path), a batch with many takes the per-query
    kernel.  Same DB, same queries through both (one call per query vs one call for all) and SQLite."""
    rng = np.random.default_rng(5)
    db = synth_db.make_db(4000, 5, 30, seed=31, lo=-22.0, hi=42.0, near_int_frac=0.5)
    sq = oracle.SqliteDB()
    for u, y in db:
        sq.add_audio(u, y)
    gpu_ctx.db_load(*synth_db.db_arrays(db))
    ys = []
    for qi in range(36):                      # every query: <= 4 integer values of max1, batch: ~60
        base = -20 + (qi * 7) % 58
        ys.append(synth_db.random_y(rng, int(rng.integers(3, 80)), lo=base, hi=base + 3.99, near_int_frac=0.6))
    ys.append(db[17][1].copy())
    ys.append(np.zeros((0, 2)))
    foff = np.zeros(len(ys) + 1, np.uint64); foff[1:] = np.cumsum([y.shape[0] for y in ys])
    ints = {int(v) for y in ys for v in np.trunc(y[:, 0])}
    assert 33 <= len(ints) <= 64, len(ints)      # coefs = 1: the batch runs on 64-bit patterns, every single on the direct table
    decoy = np.stack([np.arange(80) - 30 + 0.5, np.zeros(80)], axis=1)          # 80 more windows: the per-query kernel
    foff_d = np.concatenate([foff, [foff[-1] + 80]]).astype(np.uint64)
    for coefs, tol in ((1, 0.001), (1, 0.02), (2, 1.5)):
        batch = gpu_ctx.match(np.concatenate(ys), foff, coefs, tol)
        general = gpu_ctx.match(np.concatenate(ys + [decoy]), foff_d, coefs, tol)
        for qi, y in enumerate(ys):
            single = gpu_ctx.match(y, None, coefs, tol)[0] if y.shape[0] else batch[qi]
            assert gpu_result(single) == gpu_result(batch[qi]) == gpu_result(general[qi]), (coefs, tol, qi)
            if qi % 3 == 0 and y.shape[0]:
                assert gpu_result(batch[qi]) == sql_result(sq.search(y, coefs, tol, has_y=np.isfinite(y))), (coefs, tol, qi)
    # the table boundaries: 8 distinct windows (direct pattern table), 9..16 (hashed, 16-bit patterns),
    # 17..32 (32-bit patterns, the block's uuids in halves), 33..64 (64-bit patterns, in eighths), 65 (per-query)
    for n_int in (8, 9, 11, 12, 16, 17, 32, 33, 48, 64, 65, 70):
        y = np.stack([np.arange(n_int) - 5 + 0.2, np.zeros(n_int)], axis=1)
        h = gpu_ctx.match(y, None, 1, 0.3)[0]
        assert gpu_result(h) == sql_result(sq.search(y, 1, 0.3)), n_int


def test_hashed_pattern_tables(gpu_ctx, oracle):
    """13..32 distinct windows in a batch (recordings of very different loudness): the shared-window
    path keeps 'greatest rank per pattern' in hash tables.  (a) a batch of 24 windows against singles
    and SQLite; (b) wide windows over a DB whose audios carry ~20 000 different bit patterns: the
    batch's table (8 192 patterns) fills up and the batch is handed to the per-query kernel on the
    device -- same answers."""
    rng = np.random.default_rng(77)
    db = synth_db.make_db(6000, 4, 24, seed=91, lo=-6.0, hi=26.0, near_int_frac=0.7)
    sq = oracle.SqliteDB()
    for u, y in db:
        sq.add_audio(u, y)
    gpu_ctx.db_load(*synth_db.db_arrays(db))
    ys = []
    for qi in range(48):
        base = -4 + (qi * 5) % 22  # all queries together: integers -4 .. 19 -> 24 windows
        ys.append(synth_db.random_y(rng, int(rng.integers(2, 60)), lo=base, hi=base + 1.99, near_int_frac=0.7))
    ys.append(db[123][1].copy())
    foff = np.zeros(len(ys) + 1, np.uint64); foff[1:] = np.cumsum([y.shape[0] for y in ys])
    ints = {int(v) for y in ys for v in np.trunc(y[:, 0])}
    assert 13 <= len(ints) <= 32, len(ints)
    for tol in (0.001, 0.05):
        batch = gpu_ctx.match(np.concatenate(ys), foff, 1, tol)
        for qi, y in enumerate(ys):
            assert gpu_result(gpu_ctx.match(y, None, 1, tol)[0]) == gpu_result(batch[qi]), (tol, qi)
            if qi % 4 == 0:
                assert gpu_result(batch[qi]) == sql_result(sq.search(y, 1, tol, has_y=np.isfinite(y))), (tol, qi)
    # (b) every row lies in some window (tol 0.5 around every integer the DB uses): patterns are dense
    db2 = synth_db.make_db(20000, 6, 14, seed=92, lo=0.0, hi=19.99)
    sq2 = oracle.SqliteDB()
    for u, y in db2[:3000]:
        sq2.add_audio(u, y)
    gpu_ctx.db_load(*synth_db.db_arrays(db2))
    y_all = np.stack([np.arange(20) + 0.5, np.zeros(20)], axis=1)             # 20 windows, all occupied
    qs = [y_all, y_all[:14], db2[5][1].copy(), np.concatenate([y_all, y_all[3:9]])]
    foff = np.zeros(len(qs) + 1, np.uint64); foff[1:] = np.cumsum([y.shape[0] for y in qs])
    batch = gpu_ctx.match(np.concatenate(qs), foff, 1, 0.5)
    for qi, y in enumerate(qs):
        single = gpu_ctx.match(y, None, 1, 0.5)[0]
        assert gpu_result(single) == gpu_result(batch[qi]), qi
    # and against SQLite on the first 3 000 audios
    gpu_ctx.db_load(*synth_db.db_arrays(db2[:3000]))
    batch = gpu_ctx.match(np.concatenate(qs), foff, 1, 0.5)
    for qi, y in enumerate(qs):
        assert gpu_result(batch[qi]) == sql_result(sq2.search(y, 1, 0.5)), qi


def test_long_queries_fold_in_place(gpu_ctx, oracle):
    """Queries longer than the 1 024 frames qprep folds in shared memory (a 30 s recording is 938 frames,
    a minute is 1 875) take the in-place path; short and long queries mixed in one batch."""
    rng = np.random.default_rng(12)
    db = synth_db.make_db(500, 10, 40, seed=5)
    sq = oracle.SqliteDB()
    for u, y in db:
        sq.add_audio(u, y)
    gpu_ctx.db_load(*synth_db.db_arrays(db))
    ys = [synth_db.random_y(rng, 1875, near_int_frac=0.5), db[3][1], synth_db.random_y(rng, 1025, null_frac=0.05),
          synth_db.random_y(rng, 1024), synth_db.random_y(rng, 4000, lo=10.0, hi=30.0)]
    foff = np.zeros(len(ys) + 1, np.uint64); foff[1:] = np.cumsum([y.shape[0] for y in ys])
    for coefs, tol in ((1, 0.01), (2, 1.0)):
        hits = gpu_ctx.match(np.concatenate(ys), foff, coefs, tol)
        for y, h in zip(ys, hits):
            assert gpu_result(h) == sql_result(sq.search(y, coefs, tol, has_y=np.isfinite(y))), (coefs, tol, y.shape)


def test_very_long_queries_and_far_values(gpu_ctx, oracle):
    """A recording of more than 65 535 frames (6 min at 44.1 kHz) must not fail nor wrap a counter: a uuid
    that is hit by every frame collects more votes than a u16 holds (the per-query kernel switches to u32
    counters; the shared-window path counts in u32 anyway).  And max1 values outside the histogram qprep
    folds with (|y| >= 512: only a caller-supplied y can be) take the serial tail."""
    rng = np.random.default_rng(99)
    db = synth_db.make_db(40, 10, 20, seed=17, lo=10.0, hi=60.0, near_int_frac=0.7)
    allk = np.arange(10, 60, dtype=np.float64)
    db.append((synth.uuid_for(77_000_001), np.stack([allk + 0.0002, np.full(allk.size, 3.0)], axis=1)))   # a row next to every integer
    db.append((synth.uuid_for(77_000_002), np.array([[600.0004, 1.0], [-700.0003, 2.0]])))               # far values
    sq = oracle.SqliteDB()
    for u, y in db:
        sq.add_audio(u, y)
    gpu_ctx.db_load(*synth_db.db_arrays(db))
    n_long = 70_000
    y_long = np.stack([rng.integers(10, 60, n_long) + rng.uniform(0.0, 0.9, n_long), np.full(n_long, 3.0)], axis=1)   # 50 distinct windows
    y_far = np.array([[600.7, 0.0], [-700.2, 0.0], [15.5, 0.0], [600.1, 0.0], [2000.0, 0.0]])
    y_wide = np.stack([np.arange(-20, 80) + 0.25, np.full(100, 3.0)], axis=1)   # 100 more: the batch takes the per-query kernel (u32 counters)
    ys = [y_long, y_far, db[2][1], y_wide]
    foff = np.zeros(len(ys) + 1, np.uint64); foff[1:] = np.cumsum([y.shape[0] for y in ys])
    for coefs, tol in ((1, 0.001), (2, 0.5)):
        hits = gpu_ctx.match(np.concatenate(ys), foff, coefs, tol)
        for y, h in zip(ys, hits):
            assert gpu_result(h) == sql_result(sq.search(y, coefs, tol, has_y=np.isfinite(y))), (coefs, tol, y.shape)
    assert gpu_result(gpu_ctx.match(y_long, None, 1, 0.001)[0])[1] == n_long      # every frame votes for the all-integers audio


def test_ties_resolve_to_greatest_uuid(gpu_ctx, oracle):
    # many audios with identical rows, spread over several index blocks (> 16384 uuids)
    n = 40000
    rows = np.array([[17.0, 1.0], [18.0, 2.0], [16.0004, 3.0]])
    uu = [synth.uuid_for(50_000_000 + i) for i in range(n)]
    ub = np.stack([capi.uuid_to_bytes(u) for u in uu])
    v = np.tile(synth_db.quantize_y(rows), (n, 1))
    gpu_ctx.db_load(ub, np.arange(n + 1, dtype=np.uint64) * 3, v[:, 0], v[:, 1])
    h = gpu_ctx.match(np.array([[17.3, 0], [18.2, 0], [16.9, 0], [5.0, 0]]))[0]
    assert gpu_result(h) == (max(uu), 3, 4)
    sq = oracle.SqliteDB()
    for u in uu[:300]:
        sq.add_audio(u, rows)
    gpu_ctx.db_load(ub[:300], np.arange(301, dtype=np.uint64) * 3, v[:900, 0], v[:900, 1])
    y = np.array([[17.3, 0], [18.2, 0], [16.9, 0], [5.0, 0]])
    assert gpu_result(gpu_ctx.match(y)[0]) == sql_result(sq.search(y))


def test_add_remove_follow_sqlite(gpu_ctx, oracle):
    rng = np.random.default_rng(11)
    db = synth_db.make_db(60, 10, 30, seed=9)
    sq = oracle.SqliteDB()
    gpu_ctx.db_load(*synth_db.db_arrays([]))
    assert gpu_ctx.db_stats() == (0, 0)
    assert gpu_ctx.match(db[0][1])[0]["match_count"] == 0
    for u, y in db:
        sq.add_audio(u, y)
        v = synth_db.quantize_y(y)
        gpu_ctx.db_add(capi.uuid_to_bytes(u), v[:, 0], v[:, 1])
    queries = [db[i][1] for i in (0, 7, 33)] + [synth_db.random_y(rng, 40) for _ in range(4)]
    for y in queries:
        assert gpu_result(gpu_ctx.match(y, tolerance=0.01)[0]) == sql_result(sq.search(y, tolerance=0.01))
    for i in (7, 0, 59):
        sq.delete_audio(db[i][0])
        gpu_ctx.db_remove(capi.uuid_to_bytes(db[i][0]))
    assert gpu_ctx.db_stats() == (57, sq.count_rows())
    for y in queries:
        assert gpu_result(gpu_ctx.match(y, tolerance=0.01)[0]) == sql_result(sq.search(y, tolerance=0.01))
    with pytest.raises(capi.TirError) as e:
        gpu_ctx.db_remove(capi.uuid_to_bytes(db[7][0]))
    assert e.value.code == capi.ERR_NOTFOUND


def test_add_remove_are_incremental_on_a_large_table(oracle):
    """tir_db_add / tir_db_remove against a 1 M-fingerprint table (94 M rows): no full re-sort -- the added audios
    live in the tail index, the removed ones get tombstones -- the first search after a change costs about a
    millisecond instead of the table's sort, and the results follow SQLite (on the part of the table SQLite can
    hold: the probes below only involve audios that exist in both)."""
    import time
    import torch
    ctx = capi.Context(device=0)
    try:
        n, F = 1_000_000, 94
        g = torch.Generator(device="cuda"); g.manual_seed(5)
        uu = torch.randint(0, 256, (n, 16), dtype=torch.uint8, device="cuda", generator=g)
        v1 = torch.randint(20_100_000, 29_900_000, (n * F,), dtype=torch.int32, device="cuda", generator=g)   # never within 0.1 of an integer <= 19
        v1 = v1 - (v1 % 1_000_000) + 100_000 + (v1 % 800_000)                                                 # ... nor of any integer at all
        v2 = torch.randint(-5_000_000, 20_000_000, (n * F,), dtype=torch.int32, device="cuda", generator=g)
        ro = torch.arange(n + 1, device="cuda", dtype=torch.int64) * F
        ctx.db_load_dev(n, uu.data_ptr(), ro.data_ptr(), v1.data_ptr(), v2.data_ptr(), n * F)
        assert ctx.db_index_stats()["full_builds"] == 1
        # the big table cannot match anything at tolerance 0.01 (no max1 within 0.1 of an integer): the small DB decides
        db = synth_db.make_db(40, 10, 30, seed=31, lo=14.0, hi=19.0, near_int_frac=0.7)
        sq = oracle.SqliteDB()
        rng = np.random.default_rng(8)
        queries = [db[i][1] for i in (0, 5, 17)] + [synth_db.random_y(rng, 50, lo=14.0, hi=19.0, near_int_frac=0.6) for _ in range(5)]

        def check():
            for y in queries:
                assert gpu_result(ctx.match(y, tolerance=0.01)[0]) == sql_result(sq.search(y, tolerance=0.01))

        ctx.match(queries[0], tolerance=0.01)      # warm
        for i, (u, y) in enumerate(db):
            v = synth_db.quantize_y(y)
            sq.add_audio(u, y)
            ctx.db_add(capi.uuid_to_bytes(u), v[:, 0], v[:, 1])
            if i in (0, 1, 20):
                t0 = time.perf_counter(); ctx.match(queries[0], tolerance=0.01); dt = time.perf_counter() - t0
                assert dt < 0.05, dt                # (a full sort of 94 M rows takes ~10x that)
                check()
        check()
        st = ctx.db_index_stats()
        assert st["full_builds"] == 1 and st["tail_audios"] == 40 and st["tail_builds"] >= 3
        for i in (5, 0, 33):                        # tombstones in the tail; one in the main index too
            sq.delete_audio(db[i][0])
            ctx.db_remove(capi.uuid_to_bytes(db[i][0]))
        ctx.db_remove(uu[12345].cpu().numpy())
        t0 = time.perf_counter(); ctx.match(queries[0], tolerance=0.01); dt = time.perf_counter() - t0
        assert dt < 0.05, dt
        check()
        st = ctx.db_index_stats()
        assert st["full_builds"] == 1 and st["tombstones"] == 4 and ctx.db_stats()[0] == n + 40 - 4
        # a removed audio of the MAIN index really stops winning: give a probe that only it can match
        probe = np.stack([np.full(5, 7.0), np.zeros(5)], axis=1)
        ctx.db_add(capi.uuid_to_bytes(synth.uuid_for(88_000_001)), np.full(3, 7_000_300, np.int32), np.zeros(3, np.int32))
        assert ctx.match(probe, tolerance=0.001)[0]["match_count"] == 5
        ctx.db_remove(capi.uuid_to_bytes(synth.uuid_for(88_000_001)))
        assert ctx.match(probe, tolerance=0.001)[0]["match_count"] == 0
    finally:
        ctx.close()


def test_argument_rules(gpu_ctx):
    db = synth_db.make_db(5, 5, 5, seed=4)
    gpu_ctx.db_load(*synth_db.db_arrays(db))
    for bad in (0, 3):
        with pytest.raises(capi.TirError) as e:      # src/fp_handler.c:247 -> NULL
            gpu_ctx.match(db[0][1], coefs=bad)
        assert e.value.code == capi.ERR_ARG
    h = gpu_ctx.match(np.zeros((0, 2)), np.array([0, 0], np.uint64))[0]
    assert h["match_count"] == 0 and h["frame_count"] == 0


def test_search_end_to_end_equals_oracle_chain(gpu_ctx, oracle):
    """fp_search_fingerprint_info: PCM in, winner out.  Oracle chain = oracle extraction of the DB
    clips -> SQLite ingest -> oracle extraction of the query -> SQLite search."""
    plan = oracle.Plan()
    pcm, off = synth.make_corpus(30, 3.0, first_index=3000)
    sq = oracle.SqliteDB()
    coef, vq = gpu_ctx.extract(pcm, off)
    fo = np.concatenate([[0], np.cumsum((np.diff(off.astype(np.int64)) + 255) // 256)]).astype(np.uint64)
    uu = [synth.uuid_for(777000 + i) for i in range(30)]
    for i, u in enumerate(uu):
        _, y, _ = plan.extract(pcm[int(off[i]):int(off[i + 1])])
        sq.add_audio(u, y)
    gpu_ctx.db_load(np.stack([capi.uuid_to_bytes(u) for u in uu]), fo, vq[:, 0], vq[:, 1])
    q_idx = [0, 5, 29, 13]
    q_clips = [pcm[int(off[i]):int(off[i + 1])] for i in q_idx] + [synth.make_clip(99999, 3.0)]
    qoff = np.zeros(len(q_clips) + 1, np.uint64); qoff[1:] = np.cumsum([c.size for c in q_clips])
    for tol in (0.001, 0.05):
        hits = gpu_ctx.search(np.concatenate(q_clips), qoff, tolerance=tol)
        for c, h in zip(q_clips, hits):
            _, y, _ = plan.extract(c)
            assert gpu_result(h) == sql_result(sq.search(y, tolerance=tol, has_y=np.isfinite(y)))


def test_shard_merge_equals_unsharded(gpu_ctx, oracle):
    """DB split by uuid over S shards (S contexts on this one GPU), per-shard winners merged by
    tir_merge_hits_dev == unsharded winner, for S in 1, 2, 4, 8."""
    import torch
    rng = np.random.default_rng(21)
    db = synth_db.make_db(3000, 5, 20, seed=21)
    for i in range(30):
        db.append((synth.uuid_for(4_000_000 + i), db[i][1].copy()))      # ties across shards
    ys = [db[int(rng.integers(0, len(db)))][1] for _ in range(6)] + [synth_db.random_y(rng, 50) for _ in range(6)]
    foff = np.zeros(len(ys) + 1, np.uint64); foff[1:] = np.cumsum([y.shape[0] for y in ys])
    yall = np.concatenate(ys)
    gpu_ctx.db_load(*synth_db.db_arrays(db))
    want = gpu_ctx.match(yall, foff, tolerance=0.01)
    for S in (2, 4, 8):
        parts = [[] for _ in range(S)]
        for u, y in db:
            parts[capi.shard_of(capi.uuid_to_bytes(u), S)].append((u, y))
        gathered = []
        for part in parts:
            c = capi.Context(device=0)
            c.db_load(*synth_db.db_arrays(part))
            gathered.append(c.match(yall, foff, tolerance=0.01))
            c.close()
        g = torch.from_numpy(np.concatenate(gathered).view(np.uint8)).cuda()
        out = torch.zeros(len(ys) * 24, dtype=torch.uint8, device="cuda")
        gpu_ctx.merge_hits_dev(g.data_ptr(), S, len(ys), out.data_ptr())
        torch.cuda.synchronize()
        got = out.cpu().numpy().view(capi.HIT_DTYPE)
        assert np.array_equal(got["match_count"], want["match_count"])
        assert np.array_equal(got["uuid"], want["uuid"])
        assert np.array_equal(got["frame_count"], want["frame_count"])


def test_one_million_fingerprints_both_paths_and_brute_force(gpu_ctx):
    """BASELINE config[2] at its full size (1 M fingerprints x 94 frames, generated on the device): the
    SQLite oracle cannot ingest that in a test, so (i) the shared-window path (one batched call) and the
    per-query path (a batch widened past 64 distinct windows by a decoy query) must agree on every
    query, (ii) a brute-force restatement of the vote in torch confirms sampled queries, (iii) sharding
    the same table four ways and merging gives the same winners."""
    import torch
    dev = torch.device("cuda", 0)
    n, F, Q = 1_000_000, 94, 64
    g = torch.Generator(device=dev); g.manual_seed(123)
    uu = torch.randint(0, 256, (n, 16), dtype=torch.uint8, device=dev, generator=g)
    v1 = torch.randint(15_500_000, 18_500_000, (n * F,), dtype=torch.int32, device=dev, generator=g)
    v2 = torch.randint(-5_000_000, 20_000_000, (n * F,), dtype=torch.int32, device=dev, generator=g)
    off = torch.arange(n + 1, device=dev, dtype=torch.int64) * F
    torch.cuda.synchronize()
    gpu_ctx.db_load_dev(n, uu.data_ptr(), off.data_ptr(), v1.data_ptr(), v2.data_ptr(), n * F)
    y = np.stack([np.random.default_rng(1).uniform(15.5, 18.5, (Q, F)), np.zeros((Q, F))], axis=2)
    y[:8, :, 0] = (v1.view(n, F)[:8].cpu().numpy() * 1e-6)                     # stored entries as queries
    foff = np.arange(Q + 1, dtype=np.uint64) * F
    shared = gpu_ctx.match(y.reshape(-1, 2), foff, 1, 0.001)
    decoy = np.stack([np.arange(90) - 45.0, np.zeros(90)], axis=1)              # 89 more distinct windows -> per-query path
    foff2 = np.concatenate([foff, [foff[-1] + 90]]).astype(np.uint64)
    general = gpu_ctx.match(np.concatenate([y.reshape(-1, 2), decoy]), foff2, 1, 0.001)[:Q]
    assert np.array_equal(shared["match_count"], general["match_count"]) and np.array_equal(shared["uuid"], general["uuid"])
    assert (shared["frame_count"] == F).all() and (shared["match_count"] > 0).all()
    uub = uu.cpu().numpy()
    order = np.lexsort(uub.T[::-1]); rank_of = np.empty(n, np.int64); rank_of[order] = np.arange(n)
    rank_t = torch.from_numpy(rank_of).to(dev)
    for q in (0, 7, 8, Q - 1):
        ks, w = np.unique(np.trunc(y[q, :, 0]).astype(np.int64), return_counts=True)
        votes = torch.zeros(n, dtype=torch.int64, device=dev)
        for k, wk in zip(ks, w):
            votes += ((v1 >= int(k) * 1_000_000 - 1000) & (v1 <= int(k) * 1_000_000 + 1000)).view(n, F).any(dim=1).to(torch.int64) * int(wk)
        best = int(votes.max().item())
        cand = torch.nonzero(votes == best).flatten()
        win = int(cand[torch.argmax(rank_t[cand])].item())
        assert shared["match_count"][q] == best and bytes(shared["uuid"][q].tolist()) == bytes(uub[win].tolist())
    # four shards by uuid, merged
    shard = np.array([capi.shard_of(u, 4) for u in uub[:200000]])                # (python loop: a 200 k prefix decides the split test)
    sub_n = 200000
    ctx_all = capi.Context(device=0)
    ctx_all.db_load(uub[:sub_n], np.arange(sub_n + 1, dtype=np.uint64) * F, v1[: sub_n * F].cpu().numpy(), v2[: sub_n * F].cpu().numpy())
    want = ctx_all.match(y.reshape(-1, 2), foff, 1, 0.001)
    ctx_all.close()
    gathered = []
    for s_ in range(4):
        idx = np.nonzero(shard == s_)[0]
        rows = (idx[:, None] * F + np.arange(F)[None, :]).reshape(-1)
        c = capi.Context(device=0)
        c.db_load(uub[idx], np.arange(idx.size + 1, dtype=np.uint64) * F, v1.cpu().numpy()[rows], v2.cpu().numpy()[rows])
        gathered.append(c.match(y.reshape(-1, 2), foff, 1, 0.001))
        c.close()
    gt = torch.from_numpy(np.concatenate(gathered).view(np.uint8)).cuda()
    out = torch.zeros(Q * 24, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    gpu_ctx.merge_hits_dev(gt.data_ptr(), 4, Q, out.data_ptr())
    torch.cuda.synchronize()
    got = out.cpu().numpy().view(capi.HIT_DTYPE)
    assert np.array_equal(got["match_count"], want["match_count"]) and np.array_equal(got["uuid"], want["uuid"])


def test_steady_caller_is_served_by_graph_replays(gpu_ctx, oracle):
    """The match chain of a caller that repeats the same batch shape / buffers / parameters is captured into a CUDA
    graph (after the key was seen twice on a staging slot) and replayed with one launch per batch; every replay, and
    every call after a change of parameters or of the table, still equals SQLite."""
    rng = np.random.default_rng(21)
    db = synth_db.make_db(80, 20, 40, seed=17)
    sq = oracle.SqliteDB()
    for u, y in db:
        sq.add_audio(u, y)
    gpu_ctx.db_load(*synth_db.db_arrays(db))
    qa, qb = db[3][1], synth_db.random_y(rng, db[3][1].shape[0])
    s0 = gpu_ctx.match_graph_stats()
    for i in range(16):
        y = qa if i % 2 == 0 else qb                                   # same shape, different content: same graph
        assert gpu_result(gpu_ctx.match(y, tolerance=0.01)[0]) == sql_result(sq.search(y, tolerance=0.01)), i
    s1 = gpu_ctx.match_graph_stats()
    assert s1["graph_launches"] - s0["graph_launches"] >= 8 and 1 <= s1["graphs_built"] - s0["graphs_built"] <= 4
    for tol in (0.05, 0.01, 0.05):                                     # a parameter change: plain launches, then a new graph
        for i in range(9):
            assert gpu_result(gpu_ctx.match(qa, tolerance=tol)[0]) == sql_result(sq.search(qa, tolerance=tol))
    u_new, y_new = synth_db.make_db(1, 25, 25, seed=99)[0]
    v = synth_db.quantize_y(y_new)
    gpu_ctx.db_add(capi.uuid_to_bytes(u_new), v[:, 0], v[:, 1]); sq.add_audio(u_new, y_new)
    gpu_ctx.db_remove(capi.uuid_to_bytes(db[3][0])); sq.delete_audio(db[3][0])
    for i in range(12):                                                # tail index + tombstone: two chains per call
        for y in (qa, y_new):
            assert gpu_result(gpu_ctx.match(y, tolerance=0.01)[0]) == sql_result(sq.search(y, tolerance=0.01))
    assert gpu_ctx.match_graph_stats()["graph_launches"] > s1["graph_launches"]


def test_index_built_in_passes_equals_one_pass(monkeypatch):
    """The index build sorts the rows in passes of whole index blocks (32 B of scratch per row of a PASS, not of the
    table).  Forced to one block per pass, a 3-block table with NULLs, removed audios and empty audios must give the
    same winners as the single-pass build, query by query, on both match paths."""
    rng = np.random.default_rng(77)
    n = 40000
    nrows = rng.integers(0, 6, n)                                      # some audios have no rows at all
    off = np.concatenate([[0], np.cumsum(nrows)]).astype(np.uint64)
    y1 = rng.uniform(14.0, 19.0, int(off[-1]))
    near = rng.random(y1.size) < 0.5
    y1[near] = np.round(y1[near]) + rng.uniform(-0.012, 0.012, int(near.sum()))   # half of the rows lie next to an integer
    y = np.stack([y1, rng.uniform(-5, 20, y1.size)], axis=1)
    y[rng.random(y1.size) < 0.03, 0] = np.nan                           # NULL max1
    v = synth_db.quantize_y(y)
    ub = rng.integers(0, 256, (n, 16), dtype=np.uint8)
    queries = [np.stack([rng.uniform(14, 19, 30), rng.uniform(-5, 20, 30)], axis=1) for _ in range(6)]
    foff = np.arange(7, dtype=np.uint64) * 30
    qy = np.concatenate(queries)

    def run():
        c = capi.Context(device=0)
        try:
            c.db_load(ub, off, v[:, 0], v[:, 1])
            for a in (5, 17000, 39999):
                try:
                    c.db_remove(ub[a])
                except capi.TirError:
                    pass
            out = [c.match(qy, foff, 1, 0.01), c.match(qy, foff, 2, 0.5), c.match(qy, foff, 1, 0.3)]
            return [(o["match_count"].copy(), o["uuid"].copy()) for o in out]
        finally:
            c.close()

    one = run()
    monkeypatch.setenv("TIR_BUILD_PASS_ROWS", "1")
    many = run()
    assert any((m > 0).any() for m, _ in one)
    for (m1, u1), (m2, u2) in zip(one, many):
        assert np.array_equal(m1, m2) and np.array_equal(u1, u2)


def test_churn_compacts_the_table_and_still_follows_sqlite(gpu_ctx, oracle):
    """add / remove churn: once a quarter of the main index is dead the next search rebuilds it, and the rebuild first
    drops the removed audios from the master copy (the table does not grow without bound); every answer before,
    across and after the compaction equals SQLite, adds and removes keep working on the renumbered table."""
    rng = np.random.default_rng(33)
    db = synth_db.make_db(300, 8, 20, seed=41)
    sq = oracle.SqliteDB()
    for u, y in db:
        sq.add_audio(u, y)
    gpu_ctx.db_load(*synth_db.db_arrays(db))
    queries = [db[i][1] for i in (1, 150, 299)] + [synth_db.random_y(rng, 30) for _ in range(3)]

    def check():
        for y in queries:
            assert gpu_result(gpu_ctx.match(y, tolerance=0.01)[0]) == sql_result(sq.search(y, tolerance=0.01))

    check()
    b0 = gpu_ctx.db_index_stats()["full_builds"]
    gone = list(range(0, 300, 2))                                       # half of the table: past the tombstone budget
    for k, i in enumerate(gone):
        gpu_ctx.db_remove(capi.uuid_to_bytes(db[i][0])); sq.delete_audio(db[i][0])
        if k in (10, 60):
            check()                                                     # tombstones only
    check()                                                             # this search compacts and rebuilds
    st = gpu_ctx.db_index_stats()
    assert st["full_builds"] == b0 + 1 and st["tombstones"] == 0
    assert gpu_ctx.db_stats() == (150, sq.count_rows())
    extra = synth_db.make_db(40, 8, 20, seed=43)
    for u, y in extra:                                                  # the renumbered table takes adds and removes
        v = synth_db.quantize_y(y)
        gpu_ctx.db_add(capi.uuid_to_bytes(u), v[:, 0], v[:, 1]); sq.add_audio(u, y)
    queries.append(extra[7][1])
    check()
    for i in (1, 151, 299):
        gpu_ctx.db_remove(capi.uuid_to_bytes(db[i][0])); sq.delete_audio(db[i][0])
    gpu_ctx.db_remove(capi.uuid_to_bytes(extra[7][0])); sq.delete_audio(extra[7][0])
    check()
    with pytest.raises(capi.TirError):
        gpu_ctx.db_remove(capi.uuid_to_bytes(db[0][0]))                 # compacted away long ago: unknown
    assert gpu_ctx.db_stats() == (150 + 40 - 4, sq.count_rows())


def test_coefs2_short_queries_take_the_row_major_kernel(gpu_ctx, oracle):
    """coefs = 2 (src/fp_handler.c:318-351: a window on max2 as well, per frame).  A batch whose queries all have at most
    96 windows runs tir_match2_kernel (row-major over the frames of one max1 window, candidates united per uuid); the same
    queries batched with one long recording run the frame-major kernel; both must give SQLite's answer.  Covered: several
    rows of one audio inside one frame's window (one vote per frame and uuid), frames whose max2 falls to freq_ignore (no
    predicate on max2), NULL max2 rows, a tolerance wide enough to overflow the candidate list (the item is handed to
    the frame-major kernel), more than 32 distinct max1 windows in one query (likewise)."""
    rng = np.random.default_rng(2024)
    db = synth_db.make_db(3000, 5, 40, seed=77, near_int_frac=0.6, null_frac=0.02)
    twin = np.array([[17.0003, 5.0], [17.0004, 5.0002], [16.9998, 5.0001], [17.0001, 4.9999], [18.0002, 7.5], [18.0003, 7.5004]])
    db.append((synth.uuid_for(88_000_001), twin))
    db.append((synth.uuid_for(88_000_002), np.stack([np.arange(-10, 30) + 0.0002, np.full(40, 6.0)], axis=1)))   # a row next to 40 integers
    sq = oracle.SqliteDB()
    for u, y in db:
        sq.add_audio(u, y)
    gpu_ctx.db_load(*synth_db.db_arrays(db))
    ys = [db[int(i)][1][:96].copy() for i in rng.integers(0, 3000, 6)]
    ys += [synth_db.random_y(rng, int(n), near_int_frac=0.7, null_frac=0.05 if n % 2 else 0.0) for n in (1, 2, 31, 32, 33, 64, 95, 96)]
    ys.append(np.array([[17.4, 5.0001], [17.9, 5.0], [17.2, 4.99995], [18.1, 7.5002], [18.7, 7.5001], [16.0, 5.0]]))      # the twin audio: rows share frames
    ys.append(np.stack([np.arange(-10, 30) + 0.5, np.full(40, 6.0)], axis=1))                                          # 40 max1 windows: not for the row-major kernel
    ys.append(np.zeros((0, 2)))
    long_q = synth_db.random_y(rng, 200, near_int_frac=0.7)
    foff = np.zeros(len(ys) + 1, np.uint64); foff[1:] = np.cumsum([y.shape[0] for y in ys])
    foff_l = np.concatenate([foff, [foff[-1] + long_q.shape[0]]]).astype(np.uint64)
    for tol, lo, hi in ((0.001, -1, -1), (0.01, -1, -1), (0.5, -1, -1), (2.0, 30, 60), (0.002, 50, -1), (0.0, -1, -1)):
        short = gpu_ctx.match(np.concatenate(ys), foff, 2, tol, lo, hi)
        mixed = gpu_ctx.match(np.concatenate(ys + [long_q]), foff_l, 2, tol, lo, hi)
        for qi, y in enumerate(ys):
            assert gpu_result(short[qi]) == gpu_result(mixed[qi]), (tol, lo, hi, qi)
            if y.shape[0] and (qi % 2 == 0 or qi >= 14):
                assert gpu_result(short[qi]) == sql_result(sq.search(y, 2, tol, lo, hi, has_y=np.isfinite(y))), (tol, lo, hi, qi)
        assert gpu_result(mixed[len(ys)]) == sql_result(sq.search(long_q, 2, tol, lo, hi, has_y=np.isfinite(long_q))), (tol, lo, hi)
    h = gpu_ctx.match(ys[14], None, 2, 0.001)[0]
    assert gpu_result(h) == (synth.uuid_for(88_000_001), 5, 6)           # five frames see the twin audio, once each
