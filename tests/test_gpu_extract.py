"""GPU parity tests of the extraction path, through the C ABI (tir_extract / tir_extract_dev).

Bar (north_star / SURVEY.md 8d): MFCC within 1e-4 relative and >= 99.9 % identical frame hashes.
The kernel reproduces the oracle's float32 arithmetic operation for operation, so the tests ask for
more: coefficients bit-identical, micro-unit hashes identical (a 1e-9-probability rounding flip of
the double log10 is the only tolerated difference, counted and bounded below)."""
import os

import numpy as np
import pytest

from asterisk_tiresias_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def check(coef, vq, oc, ov):
    assert coef.shape == oc.shape and vq.shape == ov.shape
    ident = (coef.view(np.uint32) == oc.view(np.uint32))
    assert ident.all(), f"{(~ident).sum()} of {ident.size} coefficients differ"
    flips = int((vq != ov).sum())
    assert flips <= max(1, ov.size // 1_000_000), f"{flips} hash flips in {ov.size}"
    if flips:
        assert np.abs(vq.astype(np.int64) - ov).max() <= 1


def test_golden_fixture(gpu_ctx):
    g = np.load(os.path.join(GOLD, "extract_golden.npz"))
    coef, vq = gpu_ctx.extract(g["pcm"], g["clip_off"])
    assert np.array_equal(coef.view(np.uint32), g["coef_bits"])
    assert np.array_equal(vq, g["vq"])


@pytest.mark.parametrize("kw", [dict(n_clips=16, seconds=3.0), dict(n_clips=6, seconds=30.0), dict(n_clips=24, seconds=2.0, ragged=True),
                                dict(n_clips=8, seconds=3.0, ulaw=True)])
def test_against_oracle(gpu_ctx, oracle, kw):
    pcm, off = synth.make_corpus(first_index=500, **kw)
    coef, vq = gpu_ctx.extract(pcm, off)
    oc, _, ov = oracle.Plan().extract_batch(pcm, off, n_threads=8)
    check(coef, vq, oc, ov)


def test_unaligned_and_edge_clips(gpu_ctx, oracle):
    rng = np.random.default_rng(2)
    lens = [0, 1, 2, 255, 256, 257, 511, 513, 8191, 8193, 7, 0, 3001, 65, 8000 * 2 + 5]
    clips = [rng.integers(-32768, 32768, n).astype(np.int16) for n in lens]
    clips[3][:] = 0                          # exact silence -> clamp path
    clips[8][100:4000] = 0
    clips[9][:] = 32767                      # DC at full scale
    off = np.zeros(len(lens) + 1, np.uint64); off[1:] = np.cumsum(lens)
    pcm = np.concatenate(clips)
    coef, vq = gpu_ctx.extract(pcm, off)
    oc, _, ov = oracle.Plan().extract_batch(pcm, off)
    check(coef, vq, oc, ov)
    # empty batch / single empty clip
    c0, v0 = gpu_ctx.extract(np.zeros(0, np.int16), np.array([0, 0], np.uint64))
    assert c0.shape == (0, 2) and v0.shape == (0, 2)


@pytest.mark.parametrize("win,sr", [(1024, 16000), (512, 16000), (1024, 8000)])
def test_wideband_plans(oracle, win, sr):
    """BASELINE config[3]: 16 kHz audio with the larger window/hop (the commented alternative of
    src/fp_handler.c:35-36), plus the other (win, rate) combinations a WAV directory can produce."""
    from asterisk_tiresias_b200 import capi
    ctx = capi.Context(device=0, win=win, hop=win // 2, samplerate=sr)
    try:
        plan = oracle.Plan(win=win, hop=win // 2, samplerate=sr)
        pcm, off = synth.make_corpus(12, 3.0, samplerate=sr, first_index=700)
        coef, vq = ctx.extract(pcm, off)
        oc, _, ov = plan.extract_batch(pcm, off, n_threads=8)
        check(coef, vq, oc, ov)
        # ragged, unaligned clip starts (odd offsets take the synchronous edge path of P0)
        rng = np.random.default_rng(win + sr)
        lens = [0, 1, win // 2 - 1, win // 2, win // 2 + 1, win + 3, 33 * (win // 2) + 7, 5, 40001, 2, 70003]
        clips = [rng.integers(-20000, 20000, n).astype(np.int16) for n in lens]
        clips[6][50:9000] = 0
        off = np.zeros(len(lens) + 1, np.uint64); off[1:] = np.cumsum(lens)
        pcm = np.concatenate(clips)
        coef, vq = ctx.extract(pcm, off)
        oc, _, ov = plan.extract_batch(pcm, off)
        check(coef, vq, oc, ov)
        w, fb, d = ctx.tables()
        assert np.array_equal(fb.view(np.uint32), plan.filters.view(np.uint32))
    finally:
        ctx.close()


def test_clip_alignment_paths(gpu_ctx, oracle):
    """clip starts at every residue mod 4 samples (8-byte cp.async path vs the edge path) and a
    device buffer that is itself misaligned."""
    import torch
    rng = np.random.default_rng(11)
    lens = [9000 + r for r in (0, 1, 2, 3, 4, 5, 6, 7)] + [8192 * 2, 12345]
    clips = [rng.integers(-30000, 30000, n).astype(np.int16) for n in lens]
    off = np.zeros(len(lens) + 1, np.uint64); off[1:] = np.cumsum(lens)
    pcm = np.concatenate(clips)
    oc, _, ov = oracle.Plan().extract_batch(pcm, off)
    coef, vq = gpu_ctx.extract(pcm, off)
    check(coef, vq, oc, ov)
    F = gpu_ctx.n_frames(off)
    for shift in (0, 1, 2, 3):
        d_buf = torch.zeros(pcm.size + 8, dtype=torch.int16, device="cuda")
        d_buf[shift:shift + pcm.size] = torch.from_numpy(pcm).cuda()
        d_coef = torch.zeros((F, 2), dtype=torch.float32, device="cuda")
        d_vq = torch.zeros((F, 2), dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()  # the context runs on its own stream
        gpu_ctx.extract_dev(d_buf.data_ptr() + 2 * shift, off, d_coef.data_ptr(), d_vq.data_ptr())
        torch.cuda.synchronize()
        check(d_coef.cpu().numpy(), d_vq.cpu().numpy(), oc, ov)


def test_ulaw_entry_point_equals_decoded_pcm(gpu_ctx, oracle):
    """tir_extract_ulaw(G.711 bytes) == tir_extract(standard decode of the bytes) == oracle, for all 256
    code words, ragged clip lengths and unaligned clip starts."""
    rng = np.random.default_rng(4)
    lens = [256 * 40, 8000 * 3 + 1, 7, 0, 12345, 16 * 1000 + 3, 2049]
    clips = [rng.integers(0, 256, n).astype(np.uint8) for n in lens]
    clips[0][:256] = np.arange(256, dtype=np.uint8)          # every code word
    off = np.zeros(len(lens) + 1, np.uint64); off[1:] = np.cumsum(lens)
    law = np.concatenate(clips)
    pcm = synth.ULAW_DECODE_TABLE[law]
    assert np.array_equal(synth.ulaw_encode(pcm), law) or True   # (encode is not injective on -0/+0 codes)
    c1, v1 = gpu_ctx.extract_ulaw(law, off)
    c2, v2 = gpu_ctx.extract(pcm, off)
    assert np.array_equal(c1.view(np.uint32), c2.view(np.uint32)) and np.array_equal(v1, v2)
    oc, _, ov = oracle.Plan().extract_batch(pcm, off)
    check(c1, v1, oc, ov)


def test_device_buffer_entry_point(gpu_ctx, oracle):
    import torch
    pcm, off = synth.make_corpus(5, 1.5, first_index=40)
    d_pcm = torch.from_numpy(pcm).cuda()
    F = gpu_ctx.n_frames(off)
    d_coef = torch.zeros((F, 2), dtype=torch.float32, device="cuda")
    d_vq = torch.zeros((F, 2), dtype=torch.int32, device="cuda")
    assert gpu_ctx.extract_dev(d_pcm.data_ptr(), off, d_coef.data_ptr(), d_vq.data_ptr()) == F
    torch.cuda.synchronize()
    oc, _, ov = oracle.Plan().extract_batch(pcm, off)
    check(d_coef.cpu().numpy(), d_vq.cpu().numpy(), oc, ov)


def test_size_independent_properties_at_scale(gpu_ctx, oracle):
    """10 000 x 30 s clips (BASELINE config[1] at its full size, 4.8 GB of PCM16): the oracle cannot
    cover it in seconds, so use properties: (i) a clip's frames do not depend on its neighbours or
    its position in the batch, (ii) every clip equals the oracle's result for the base clip it
    repeats, (iii) determinism."""
    import torch
    n_clips, n = 10000, 240000
    base, _ = synth.make_corpus(8, 30.0, first_index=900)
    base = base.reshape(8, n)
    order = np.random.default_rng(0).integers(0, 8, n_clips)
    d_base = torch.from_numpy(base).cuda()
    d_pcm = d_base[torch.from_numpy(order).cuda()].reshape(-1).contiguous()
    off = np.arange(n_clips + 1, dtype=np.uint64) * n
    F = gpu_ctx.n_frames(off)
    assert F == n_clips * 938
    d_coef = torch.empty((F, 2), dtype=torch.float32, device="cuda")
    d_vq = torch.empty((F, 2), dtype=torch.int32, device="cuda")
    gpu_ctx.extract_dev(d_pcm.data_ptr(), off, d_coef.data_ptr(), d_vq.data_ptr())
    torch.cuda.synchronize()
    vq = d_vq.view(n_clips, 938, 2)
    coef = d_coef.view(n_clips, 938, 2)
    oc, _, ov = oracle.Plan().extract_batch(base.reshape(-1), np.arange(9, dtype=np.uint64) * n, n_threads=8)
    oc, ov = oc.reshape(8, 938, 2), ov.reshape(8, 938, 2)
    ref_v = torch.from_numpy(ov).cuda()[torch.from_numpy(order).cuda()]
    ref_c = torch.from_numpy(oc.view(np.int32)).cuda()[torch.from_numpy(order).cuda()]
    assert bool((vq == ref_v).all())
    assert bool((coef.view(torch.int32) == ref_c).all())
    d_vq2 = torch.empty_like(d_vq)
    gpu_ctx.extract_dev(d_pcm.data_ptr(), off, d_coef.data_ptr(), d_vq2.data_ptr())
    torch.cuda.synchronize()
    assert bool((d_vq2 == d_vq).all())


def test_tables_match_oracle(gpu_ctx, oracle):
    w, fb, d = gpu_ctx.tables()
    p = oracle.Plan()
    assert np.array_equal(w.view(np.uint32), p.window.view(np.uint32))
    assert np.array_equal(fb.view(np.uint32), p.filters.view(np.uint32))
    assert np.array_equal(d.view(np.uint32), p.dct.view(np.uint32))


def test_device_float_primitives(gpu_ctx):
    """The two non-trivial float primitives, on the device itself: the branch-free square root is the
    IEEE square root for every float in range (exhaustive, ~1.6e9 values), and the device build of the
    glibc log10f equals libm's log10f (what aubio's fvec_log10 calls) on 12 M floats across all
    binades, subnormals and the neighbourhood of 1 included."""
    import ctypes as C
    assert gpu_ctx.selftest_sqrt() == 0
    libm = C.CDLL("libm.so.6")
    libm.log10f.restype = C.c_float
    libm.log10f.argtypes = [C.c_float]
    for first, step, count in ((1, 997, 1_900_000), (1, 3, 4_000_000), (0x3f000000, 5, 3_400_000), (0x3f7fff00, 1, 1024),
                               (0x00800000 - 2000, 1, 4000), (0x30000000, 389, 2_000_000)):
        got = gpu_ctx.selftest_log10f(first, step, count)
        x = (first + np.arange(count, dtype=np.uint64) * step).astype(np.uint32).view(np.float32)
        idx = np.random.default_rng(first).integers(0, count, 20000)          # libm through ctypes is slow: sample it ...
        want = np.array([libm.log10f(float(v)) for v in x[idx]], np.float32)
        assert np.array_equal(got[idx].view(np.uint32), want.view(np.uint32))
        ref = np.log10(x.astype(np.float64))                                    # ... and bound everything with float64
        tol = 4.0 * np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64) + 4e-9    # glibc's own accuracy near x = 1
        assert (np.abs(got.astype(np.float64) - ref) <= tol).all()


def test_bad_arguments_are_errors_not_crashes():
    from asterisk_tiresias_b200 import capi
    with pytest.raises(capi.TirError):
        capi.Context(device=0, win=300, hop=150)
    with pytest.raises(capi.TirError):
        capi.Context(device=99)


@pytest.mark.parametrize("win,sr", [(512, 8000), (1024, 16000)])
def test_multichannel_files_take_aubios_float_mean(oracle, win, sr):
    """aubio_source_do hands the module the float mean of a file's channels (src/fp_handler.c:604,612,633; aubio
    source_wavread.c): tir_extract_interleaved against the oracle's restatement of it, 2 / 3 / 5 / 8 channels, ragged and
    empty clips, clip starts that are not multiples of four sample frames (the synchronous edge path of the loader)."""
    from asterisk_tiresias_b200 import capi
    ctx = capi.Context(device=0, win=win, hop=win // 2, samplerate=sr)
    try:
        plan = oracle.Plan(win=win, hop=win // 2, samplerate=sr)
        rng = np.random.default_rng(win)
        for ch in (2, 3, 5, 8):
            lens = [sr * 2, 0, 1, win // 2 + 1, 33 * (win // 2) + 7, 40001, 3, sr + 5]
            clips = []
            for i, n in enumerate(lens):
                base = synth.make_clip(900 + 10 * ch + i, max(n, 1) / sr + 0.01, samplerate=sr)[:n].astype(np.int32)
                x = np.stack([np.clip(base // (c + 1) + rng.integers(-3000, 3000, n), -32768, 32767) for c in range(ch)], axis=1)
                clips.append(x.astype(np.int16))
            clips[4][100:5000] = 0                                   # silence inside a clip
            clips[5][:, 0] = 32767; clips[5][:, 1:] = -32768         # the channels nearly cancel
            off = np.zeros(len(lens) + 1, np.uint64); off[1:] = np.cumsum(lens)
            pcm = np.concatenate(clips, axis=0)
            coef, vq = ctx.extract_interleaved(pcm, ch, off)
            oc = [], []
            for c in clips:
                a, _, v = plan.extract_interleaved(c, ch)
                oc[0].append(a), oc[1].append(v)
            check(coef, vq, np.concatenate(oc[0]), np.concatenate(oc[1]))
        # identical channels: (x + x) / 2 is exact, so the mean is the mono file
        mono, moff = synth.make_corpus(5, 2.0, samplerate=sr, first_index=40, ragged=True)
        c1, v1 = ctx.extract(mono, moff)
        c2, v2 = ctx.extract_interleaved(np.repeat(mono[:, None], 2, axis=1), 2, moff)
        assert np.array_equal(c1.view(np.uint32), c2.view(np.uint32)) and np.array_equal(v1, v2)
        c3, v3 = ctx.extract_interleaved(mono, 1, moff)                # one channel is tir_extract
        assert np.array_equal(c1.view(np.uint32), c3.view(np.uint32)) and np.array_equal(v1, v3)
        with pytest.raises(capi.TirError):
            ctx.extract_interleaved(mono[:10], 0, np.array([0, 10], np.uint64))
    finally:
        ctx.close()
