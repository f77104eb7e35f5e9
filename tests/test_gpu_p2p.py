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
