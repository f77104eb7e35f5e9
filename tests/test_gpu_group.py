"""Several contexts in one process (tir_group): the table sharded by uuid over the devices, the
coefficients and the per-shard winners moved with peer copies.  Results must equal one context
holding the whole table.  Runs on every visible GPU, and with several contexts on ONE device (the
sharding, the fan-out and the merge are the same code), so it is covered on a single-GPU box too."""
import numpy as np
import pytest

from asterisk_tiresias_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _same(a, b):
    return (np.array_equal(a["match_count"], b["match_count"]) and np.array_equal(a["frame_count"], b["frame_count"])
            and np.array_equal(a["uuid"][a["match_count"] > 0], b["uuid"][b["match_count"] > 0]))


@pytest.mark.parametrize("layout", ["one-device-x3", "all-devices"])
def test_group_search_equals_single_context(layout):
    import torch
    n_dev = torch.cuda.device_count()
    devices = [0, 0, 0] if layout == "one-device-x3" else list(range(n_dev))
    if layout == "all-devices" and n_dev < 2:
        pytest.skip("one GPU visible")
    pcm, off = synth.make_corpus(400, 2.0, first_index=20000, ragged=True)
    ref = capi.Context(device=0)
    grp = capi.Group(devices)
    try:
        coef, vq = ref.extract(pcm, off)
        fo = np.concatenate([[0], np.cumsum((np.diff(off.astype(np.int64)) + 255) // 256)]).astype(np.uint64)
        uu = np.stack([capi.uuid_to_bytes(synth.uuid_for(930000 + i)) for i in range(400)])
        ref.db_load(uu, fo, vq[:, 0], vq[:, 1])
        grp.db_load(uu, fo, vq[:, 0], vq[:, 1])
        assert grp.db_stats() == ref.db_stats()
        qi = list(range(0, 400, 7))
        q_clips = [pcm[int(off[i]):int(off[i + 1])] for i in qi] + [synth.make_clip(777 + j, 1.7) for j in range(9)] + [np.zeros(0, np.int16)]
        qoff = np.zeros(len(q_clips) + 1, np.uint64); qoff[1:] = np.cumsum([c.size for c in q_clips])
        qpcm = np.concatenate(q_clips)
        for coefs, tol in ((1, 0.001), (1, 0.05), (2, 0.8)):
            assert _same(grp.search(qpcm, qoff, coefs, tol), ref.search(qpcm, qoff, coefs, tol)), (coefs, tol)
        # insert / delete go to the owning shard
        extra = synth.make_clip(424243, 2.5)
        ec, ev = ref.extract(extra)
        eu = capi.uuid_to_bytes(synth.uuid_for(999999))
        for t in (ref, grp):
            t.db_add(eu, ev[:, 0], ev[:, 1])
            t.db_remove(uu[14])
        assert grp.db_stats() == ref.db_stats()
        q2 = np.concatenate([extra, q_clips[2]]); q2off = np.array([0, extra.size, extra.size + q_clips[2].size], np.uint64)
        assert _same(grp.search(q2, q2off, 1, 0.05), ref.search(q2, q2off, 1, 0.05))
        with pytest.raises(capi.TirError):
            grp.search(qpcm, qoff, 3, 0.001)
        st = grp.stats()
        if layout == "all-devices":
            assert st["fused"] >= 4 and st["copy_path"] == 0   # one device per shard: coefficients + winners cross NVLink inside the kernels
        else:
            assert st["copy_path"] >= 4 and st["fused"] == 0   # shards sharing a device never wait for one another: copy path
        # config[4] on a group: concurrent channels through the group's batcher equal lone searches
        import threading
        want = ref.search(qpcm, qoff, 1, 0.05)
        grp.batcher_start(max_batch=32, max_wait_us=2000)
        got, errs = [None] * len(q_clips), []

        def channel(i):
            try:
                got[i] = grp.search_one(q_clips[i], 1, 0.05)
            except Exception as ex:  # noqa: BLE001
                errs.append(ex)
        th = [threading.Thread(target=channel, args=(i,)) for i in range(len(q_clips))]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errs, errs
        for i, g_ in enumerate(got):
            assert g_["match_count"] == want[i]["match_count"] and g_["frame_count"] == want[i]["frame_count"]
            assert g_["match_count"] == 0 or bytes(g_["uuid"]) == bytes(want[i]["uuid"])
        n_req, n_b, mx = grp.batcher_stats()
        assert n_req == len(q_clips) and n_b < n_req
        grp.batcher_stop()
    finally:
        grp.close()
        ref.close()
