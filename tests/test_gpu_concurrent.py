"""BASELINE config[4]: many dialplan channels searching at once.  Concurrent tir_search_one() callers
(one host thread per channel, as the PBX does, src/application_handler.c:180) through the batcher
must get exactly what a lone tir_search of the same recording returns, and the streaming entry
points must equal the one-shot call."""
import threading

import numpy as np
import pytest

from asterisk_tiresias_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _db_and_clips(ctx, n_db=200, n_q=96):
    pcm, off = synth.make_corpus(n_db, 3.0, first_index=12000)
    coef, vq = ctx.extract(pcm, off)
    fo = np.concatenate([[0], np.cumsum((np.diff(off.astype(np.int64)) + 255) // 256)]).astype(np.uint64)
    uu = np.stack([capi.uuid_to_bytes(synth.uuid_for(910000 + i)) for i in range(n_db)])
    ctx.db_load(uu, fo, vq[:, 0], vq[:, 1])
    clips = []
    for i in range(n_q):
        if i % 3 == 0:
            clips.append(pcm[int(off[i]):int(off[i + 1])].copy())          # a stored recording
        elif i % 3 == 1:
            clips.append(synth.make_clip(50000 + i, 1.0 + (i % 5) * 0.7))    # unrelated, ragged lengths
        else:
            c = pcm[int(off[i]):int(off[i + 1])].astype(np.int32) + np.random.default_rng(i).integers(-3, 4, int(off[i + 1] - off[i]))
            clips.append(np.clip(c, -32768, 32767).astype(np.int16))         # noisy copy
    return clips


def test_concurrent_callers_equal_single_searches():
    ctx = capi.Context(device=0)
    try:
        clips = _db_and_clips(ctx)
        params = [(1, 0.001, -1, -1), (1, 0.05, -1, -1), (2, 0.5, -1, -1), (1, 0.05, 40, 70)]
        want = [ctx.search(c, None, *params[i % len(params)])[0] for i, c in enumerate(clips)]
        ctx.batcher_start(max_batch=64, max_wait_us=2000)
        got = [None] * len(clips)
        errs = []

        def channel(i):
            try:
                got[i] = ctx.search_one(clips[i], *params[i % len(params)])
            except Exception as ex:  # noqa: BLE001
                errs.append(ex)

        for _ in range(3):                                   # three waves of 96 simultaneous channels
            th = [threading.Thread(target=channel, args=(i,)) for i in range(len(clips))]
            [t.start() for t in th]
            [t.join() for t in th]
            assert not errs, errs
            for w, g in zip(want, got):
                assert g["match_count"] == w["match_count"] and g["frame_count"] == w["frame_count"]
                assert bytes(g["uuid"]) == bytes(w["uuid"])
        n_req, n_batches, max_seen = ctx.batcher_stats()
        assert n_req == 3 * len(clips) and n_batches < n_req and max_seen > 1     # requests really were batched
        with pytest.raises(capi.TirError):                   # argument rule of src/fp_handler.c:247 survives the queue
            ctx.search_one(clips[0], coefs=3)
        ctx.batcher_stop()
        h = ctx.search_one(clips[0])                         # no batcher: a batch of one
        assert h["match_count"] == want[0]["match_count"]
    finally:
        ctx.close()


def test_streaming_equals_one_shot():
    ctx = capi.Context(device=0)
    try:
        clips = _db_and_clips(ctx, n_db=60, n_q=12)
        ctx.batcher_start(max_batch=16, max_wait_us=500)
        for c in clips:
            s = ctx.stream()
            for a in range(0, c.size, 160):                  # 20 ms slinear frames at 8 kHz
                s.feed(c[a:a + 160])
            assert s.samples == c.size
            g, w = s.finish(tolerance=0.01), ctx.search(c, None, 1, 0.01)[0]
            s.close()
            assert g["match_count"] == w["match_count"] and g["frame_count"] == w["frame_count"] and bytes(g["uuid"]) == bytes(w["uuid"])
        s = ctx.stream()                                      # empty recording: no frames, NOTFOUND
        assert s.finish()["match_count"] == 0
        s.close()
    finally:
        ctx.close()
