"""The extraction oracle against independent second opinions (numpy float64 FFT, scipy DCT,
closed forms) -- SURVEY.md section 4: the reference ships no tests, so these pin the restatement."""
import numpy as np
import pytest
import scipy.fft
import scipy.signal

from asterisk_tiresias_b200 import synth


@pytest.mark.parametrize("win,sr", [(512, 8000), (1024, 16000)])
def test_window_is_periodic_hann(oracle, win, sr):
    p = oracle.Plan(win=win, hop=win // 2, samplerate=sr)
    ref = scipy.signal.get_window("hann", win, fftbins=True)
    assert np.abs(p.window - ref).max() < 2e-7
    assert p.window[0] == 0.0 and p.window[win // 2] == 1.0


def test_slaney_band_edges_and_filterbank_shape(oracle):
    p = oracle.Plan(samplerate=8000)
    e = p.band_edges
    assert e.shape == (42,)
    assert abs(e[0] - 133.3333) < 1e-3 and abs(e[12] - 933.3333) < 1e-3
    assert abs(e[41] - 6853.8) < 0.2          # top edge of the 40th triangle
    fb = p.filters
    assert fb.shape == (40, 257)
    assert (fb >= 0).all()
    assert int((fb != 0).sum()) == 490         # SURVEY.md 8a probe
    assert [i for i in range(40) if not fb[i].any()] == [34, 35, 36, 37, 38, 39]  # edges above Nyquist
    assert not fb[:, 0].any() and not fb[:, 256].any()
    p16 = oracle.Plan(win=1024, hop=512, samplerate=16000)
    assert all(p16.filters[i].any() for i in range(40))
    # unit-area triangles: sum(weights) * bin_hz ~ 1 for filters well inside the band
    area = p16.filters.sum(axis=1) * (16000 / 1024)
    assert np.abs(area[5:35] - 1.0).max() < 0.1


def test_dct_matches_scipy_ortho(oracle):
    p = oracle.Plan()
    eye = np.eye(40)
    ref = scipy.fft.dct(eye, type=2, norm="ortho", axis=0)[:2]
    assert np.abs(p.dct - ref).max() < 1e-6


@pytest.mark.parametrize("win", [512, 1024])
def test_tir_fft_against_numpy(oracle, win):
    p = oracle.Plan(win=win, hop=win // 2, samplerate=8000 if win == 512 else 16000)
    rng = np.random.default_rng(3)
    for scale in (1.0, 1e-3, 30.0):
        x = (rng.normal(size=win) * scale).astype(np.float32)
        re, im = p.rfft(x)
        X = np.fft.rfft(x.astype(np.float64))
        err = np.abs((re.astype(np.float64) + 1j * im) - X).max() / np.abs(X).max()
        assert err < 5e-7
    # impulse and DC known answers
    x = np.zeros(win, np.float32); x[0] = 1
    re, im = p.rfft(x)
    assert np.array_equal(re, np.ones(win // 2 + 1, np.float32)) and not im.any()
    re, im = p.rfft(np.ones(win, np.float32))
    assert re[0] == win and np.abs(re[1:]).max() < 1e-4


def test_pvoc_norm_is_shift_invariant_magnitude(oracle):
    p = oracle.Plan()
    rng = np.random.default_rng(4)
    data = rng.normal(size=512).astype(np.float32)
    norm = p.pvoc_norm(data)
    ref = np.abs(np.fft.rfft(data.astype(np.float64) * p.window.astype(np.float64)))
    assert np.abs(norm - ref).max() / ref.max() < 5e-7


@pytest.mark.parametrize("kind", ["tone", "noise", "chirp", "composite"])
def test_full_pipeline_against_float64(oracle, kind):
    p = oracle.Plan()
    pcm = synth.make_clip(11, 3.0, kind=kind, silence=False)
    coef, y, vq = p.extract(pcm)
    F = coef.shape[0]
    assert F == -(-pcm.size // 256) == 94
    x = pcm.astype(np.float64) / 32768
    xx = np.concatenate([np.zeros(256), x, np.zeros(512)])
    frames = np.stack([xx[t * 256:t * 256 + 512] for t in range(F)])
    w = 0.5 * (1 - np.cos(2 * np.pi * np.arange(512) / 512))
    spec = np.abs(np.fft.rfft(frames * w, axis=1))
    mel = spec @ p.filters.astype(np.float64).T
    c64 = np.log10(np.maximum(mel, 2e-42)) @ p.dct.astype(np.float64).T
    rel = np.abs(c64 - coef) / np.maximum(np.abs(c64), 1e-3 * np.abs(c64).max(axis=0))
    # SURVEY.md 8d parity gate: MFCC within 1e-4 relative.  A pure tone is the hard case: far
    # mel bands sit at the float32 FFT noise floor, which is the reference's own arithmetic.
    assert rel.max() < (1e-4 if kind != "tone" else 5e-4)
    y64 = 10 * np.log10(np.abs(coef.astype(np.float64)))
    assert np.allclose(y, y64, rtol=1e-14, atol=0)
    assert np.array_equal(vq, np.round(y * 1e6).astype(np.int32))


def test_frame_count_and_zero_history(oracle):
    p = oracle.Plan()
    assert p.n_frames(0) == 0 and p.n_frames(1) == 1 and p.n_frames(256) == 1 and p.n_frames(257) == 2
    assert p.n_frames(240000) == 938 and p.n_frames(24000) == 94
    pcm = synth.make_clip(5, 0.2, kind="noise")
    c_full, _, _ = p.extract(pcm)
    # frame t only depends on samples [(t-1)*hop, (t+1)*hop): re-extracting a suffix that starts on
    # a hop boundary with one hop of history reproduces the later frames exactly
    c_tail, _, _ = p.extract(pcm[256 * 2:])
    assert np.array_equal(c_full[3:], c_tail[1:])


def test_silence_and_null(oracle):
    p = oracle.Plan()
    coef, y, vq = p.extract(np.zeros(1000, np.int16))
    # every band clamps to 2e-42 -> log10 = -41.69897; c0 = 40 * that / sqrt(40)
    assert np.allclose(coef[:, 0], np.log10(np.float32(2e-42)) * 40 / np.sqrt(40), rtol=1e-5)
    assert (vq[:, 0] != oracle.NULL_V).all()
    assert oracle.quantize(float("inf")) == oracle.NULL_V and oracle.quantize(float("nan")) == oracle.NULL_V
    assert oracle.quantize(17.1234565) in (17123456, 17123457)
    assert oracle.quantize(-3.0000004) == -3000000 and oracle.quantize(0.0) == 0


def test_batch_equals_single(oracle):
    p = oracle.Plan()
    pcm, off = synth.make_corpus(6, 1.0, ragged=True)
    cb, yb, vb = p.extract_batch(pcm, off, n_threads=3)
    parts = [p.extract(pcm[int(off[i]):int(off[i + 1])]) for i in range(6)]
    assert np.array_equal(cb, np.concatenate([x[0] for x in parts]))
    assert np.array_equal(vb, np.concatenate([x[2] for x in parts]))


def test_interleaved_channels_are_averaged_in_float32_like_aubios_source(oracle):
    """aubio source_wavread.c: every channel sample * (1/32768) in float, summed in channel order from 0, divided by the
    channel count -- restated here in numpy float32 and fed to the mono oracle's stages; one channel is the mono path."""
    p = oracle.Plan()
    rng = np.random.default_rng(11)
    for ch in (1, 2, 3, 6):
        x = rng.integers(-32768, 32768, (3000 + ch, ch)).astype(np.int16)
        coef, y, vq = p.extract_interleaved(x, ch)
        acc = np.zeros(x.shape[0], np.float32)
        for c in range(ch):
            acc = (acc + x[:, c].astype(np.float32) / np.float32(32768)).astype(np.float32)
        mono = (acc / np.float32(ch)).astype(np.float32)
        data = np.zeros(512, np.float32)
        for t in range(coef.shape[0]):
            hop = np.zeros(256, np.float32)
            seg = mono[t * 256:(t + 1) * 256]
            hop[:seg.size] = seg
            data = np.concatenate([data[256:], hop])
            _, c2 = p.mfcc(p.pvoc_norm(data))
            assert np.array_equal(c2.view(np.uint32), coef[t].view(np.uint32)), (ch, t)
    mono16 = synth.make_clip(3, 1.0)
    a = p.extract(mono16)
    b = p.extract_interleaved(mono16, 1)
    c = p.extract_interleaved(np.repeat(mono16[:, None], 2, axis=1), 2)    # (x + x) / 2 is exact
    for u, v, w in zip(a, b, c):
        assert np.array_equal(u, v, equal_nan=True) and np.array_equal(u, w, equal_nan=True)
