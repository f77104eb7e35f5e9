"""The cross-GPU step of the sharded match without a collective library (csrc/tir_p2p.cu): every rank
stores its winners into every peer's gather buffer over peer memory and folds the candidates once the
peers' flags have arrived.  Here: three contexts of ONE process (tir_p2p_connect_local; on one device
when the box has one GPU, spread over the devices otherwise) against one context holding the whole table."""
import numpy as np
import pytest

from asterisk_tiresias_b200 import capi, synth, synth_db

pytestmark = pytest.mark.gpu


def test_p2p_exchange_equals_unsharded():
    import torch
    world = 3
    n_dev = torch.cuda.device_count()
    rng = np.random.default_rng(21)
    db = synth_db.make_db(900, 5, 30, seed=71, lo=14.0, hi=19.0, near_int_frac=0.6)
    for i in range(20):
        db.append((synth.uuid_for(5_000_000 + i), db[i][1].copy()))   # ties across shards: copies of an audio under other uuids
    uu, row_off, v1, v2 = synth_db.db_arrays(db)
    uu = np.asarray(uu, np.uint8).reshape(-1, 16)
    full = capi.Context(device=0)
    full.db_load(uu, row_off, v1, v2)
    ctxs = [capi.Context(device=r % n_dev) for r in range(world)]
    shard = np.array([capi.shard_of(uu[a], world) for a in range(uu.shape[0])])
    for r, c in enumerate(ctxs):
        idx = np.nonzero(shard == r)[0]
        ro = np.zeros(idx.size + 1, np.uint64)
        rows = [np.arange(int(row_off[a]), int(row_off[a + 1])) for a in idx]
        ro[1:] = np.cumsum([x.size for x in rows])
        sel = np.concatenate(rows) if rows else np.zeros(0, np.int64)
        c.db_load(uu[idx], ro, v1[sel], v2[sel])
    Q = 40
    ys = [synth_db.random_y(rng, int(rng.integers(1, 50)), lo=14.0, hi=19.0, near_int_frac=0.6) for _ in range(Q - 2)]
    ys += [db[5][1].copy(), np.zeros((0, 2))]
    foff = np.zeros(Q + 1, np.uint64); foff[1:] = np.cumsum([y.shape[0] for y in ys])
    y_all = np.concatenate(ys)
    coef = np.power(10.0, y_all / 10.0).astype(np.float32)
    p2ps = [capi.P2P(c, r, world, 64) for r, c in enumerate(ctxs)]
    for p in p2ps:
        p.connect_local(p2ps)
    for coefs, tol in ((1, 0.01), (2, 1.0), (1, 0.001)):      # three batches: both gather buffers, both paths
        d_coef = [torch.from_numpy(coef).to(f"cuda:{r % n_dev}") for r in range(world)]
        d_fin = [torch.zeros(Q * 24, dtype=torch.uint8, device=f"cuda:{r % n_dev}") for r in range(world)]
        d_c0 = torch.from_numpy(coef).to("cuda:0")
        d_ref = torch.zeros(Q * 24, dtype=torch.uint8, device="cuda:0")
        full.match_dev(d_c0.data_ptr(), foff, d_ref.data_ptr(), coefs, tol)
        for r, p in enumerate(p2ps):                           # enqueue all ranks, then wait: SPMD
            p.match_dev(d_coef[r].data_ptr(), foff, d_fin[r].data_ptr(), coefs, tol)
        torch.cuda.synchronize()
        for d in range(n_dev):
            torch.cuda.synchronize(d)
        ref = d_ref.cpu().numpy().view(capi.HIT_DTYPE)
        assert (ref["match_count"] > 0).sum() > 3
        for r, p in enumerate(p2ps):
            assert p.error() == 0
            got = d_fin[r].cpu().numpy().view(capi.HIT_DTYPE)
            assert np.array_equal(got["match_count"], ref["match_count"]), (coefs, tol, r)
            assert np.array_equal(got["uuid"], ref["uuid"]) and np.array_equal(got["frame_count"], ref["frame_count"])
    for p in p2ps:
        p.close()


def _shard_load(ctxs, uu, row_off, v1, v2):
    world = len(ctxs)
    shard = np.array([capi.shard_of(uu[a], world) for a in range(uu.shape[0])])
    for r, c in enumerate(ctxs):
        idx = np.nonzero(shard == r)[0]
        ro = np.zeros(idx.size + 1, np.uint64)
        rows = [np.arange(int(row_off[a]), int(row_off[a + 1])) for a in idx]
        ro[1:] = np.cumsum([x.size for x in rows])
        sel = np.concatenate(rows) if rows else np.zeros(0, np.int64)
        c.db_load(uu[idx], ro, v1[sel], v2[sel])


def test_p2p_sharded_search_equals_unsharded():
    """tir_p2p_search: every rank uploads + extracts only its slice of the batch's clips, the coefficients
    cross the ranks inside the extraction kernel (peer stores + flag), every rank matches all queries against
    its shard and folds the winners in the last CTA of its match chain.  Must equal tir_search on one context
    holding the whole table -- for uneven slices (a rank without a clip included)."""
    import torch
    world = 3
    n_dev = torch.cuda.device_count()
    pcm, off = synth.make_corpus(90, 2.0, first_index=31000, ragged=True)
    full = capi.Context(device=0)
    coef, vq = full.extract(pcm, off)
    fo = np.concatenate([[0], np.cumsum((np.diff(off.astype(np.int64)) + 255) // 256)]).astype(np.uint64)
    uu = np.stack([capi.uuid_to_bytes(synth.uuid_for(940000 + i)) for i in range(90)])
    full.db_load(uu, fo, vq[:, 0], vq[:, 1])
    ctxs = [capi.Context(device=r % n_dev) for r in range(world)]
    _shard_load(ctxs, uu, fo, vq[:, 0], vq[:, 1])
    q_clips = [pcm[int(off[i]):int(off[i + 1])] for i in range(0, 90, 5)] + [synth.make_clip(555 + j, 1.3) for j in range(7)] + [np.zeros(0, np.int16)]
    Q = len(q_clips)
    frames = np.array([(c.size + 255) // 256 for c in q_clips], np.uint64)
    all_foff = np.concatenate([[0], np.cumsum(frames)]).astype(np.uint64)
    p2ps = [capi.P2P(c, r, world, 64, max_frames=int(all_foff[-1]) + 8) for r, c in enumerate(ctxs)]
    for p in p2ps:
        p.connect_local(p2ps)
    qoff = np.zeros(Q + 1, np.uint64); qoff[1:] = np.cumsum([c.size for c in q_clips])
    qpcm = np.concatenate(q_clips)
    for p in p2ps:      # one host thread drives all ranks here: no allocation may happen between their enqueues
        p.reserve(qpcm.size)
    for cuts, coefs, tol in (((0, 9, 20, Q), 1, 0.05), ((0, 0, Q - 3, Q), 2, 0.8), ((0, 13, 13, Q), 1, 0.001)):
        ref = full.search(qpcm, qoff, coefs, tol)
        assert (ref["match_count"] > 0).sum() > 3
        d_fin = [torch.zeros(Q * 24, dtype=torch.uint8, device=f"cuda:{r % n_dev}") for r in range(world)]
        for r, p in enumerate(p2ps):    # enqueue all ranks (device results only: no host wait inside), then synchronise
            a, b = cuts[r], cuts[r + 1]
            sl = qpcm[int(qoff[a]):int(qoff[b])]
            p.search(sl, qoff[a:b + 1] - qoff[a], a, all_foff, coefs, tol, d_final=d_fin[r].data_ptr(), want_hits=False)
        for d in range(n_dev):
            torch.cuda.synchronize(d)
        for r, p in enumerate(p2ps):
            assert p.error() == 0
            got = d_fin[r].cpu().numpy().view(capi.HIT_DTYPE)
            assert np.array_equal(got["match_count"], ref["match_count"]), (cuts, coefs, tol, r)
            assert np.array_equal(got["frame_count"], ref["frame_count"])
            assert np.array_equal(got["uuid"][got["match_count"] > 0], ref["uuid"][ref["match_count"] > 0])
    for p in p2ps:
        p.close()
    for c in ctxs + [full]:
        c.close()


def test_p2p_ipc_two_processes():
    """One process per GPU, regions opened through CUDA IPC handles (tir_p2p_connect): 2 ranks under
    torch.distributed.run, each checking the sharded match and the sharded search against a full-table
    context of its own.  Needs two visible GPUs."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU visible")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(root, "tests", "mp", "p2p_ipc_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("p2p-ipc-ok") == 2, r.stdout[-2000:]
