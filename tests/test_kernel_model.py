"""CPU checks of the product's host-side pieces: plan tables, the kernel's per-thread phase
functions run on the host (tests/emul -- test only), the glibc-exact log10f and the "%f"
quantiser.  Everything is compared bit for bit with the oracle / libc."""
import ctypes as C

import numpy as np
import pytest

from asterisk_tiresias_b200 import synth


@pytest.mark.parametrize("win,sr", [(512, 8000), (512, 16000), (1024, 16000), (512, 44100)])
def test_plan_tables_bit_identical_to_oracle(oracle, emul, win, sr):
    p = oracle.Plan(win=win, hop=win // 2, samplerate=sr)
    w = np.empty(win, np.float32); fb = np.empty((40, win // 2 + 1), np.float32); d = np.empty((2, 40), np.float32)
    assert emul.emul_tables(win, win // 2, sr, w.ctypes.data, fb.ctypes.data, d.ctypes.data) == 0
    assert np.array_equal(w.view(np.uint32), p.window.view(np.uint32))
    assert np.array_equal(fb.view(np.uint32), p.filters.view(np.uint32))
    assert np.array_equal(d.view(np.uint32), p.dct.view(np.uint32))


CASES = [(0, "tone", 3.0), (1, "noise", 3.0), (2, "chirp", 3.0), (3, "composite", 2.77), (4, "silence", 1.0),
         (5, None, 0.01), (6, None, 0.5), (7, None, 4.1)]


@pytest.mark.parametrize("idx,kind,sec", CASES)
def test_kernel_phases_on_host_match_oracle(oracle, emul, idx, kind, sec):
    p = oracle.Plan()
    pcm = synth.make_clip(idx, sec, kind=kind, ulaw=(idx % 2 == 1))
    co, y, vq = p.extract(pcm)
    F = co.shape[0]
    c2 = np.zeros((F, 2), np.float32); v2 = np.zeros((F, 2), np.int32)
    assert emul.emul_extract(pcm.ctypes.data, pcm.size, 8000, c2.ctypes.data, v2.ctypes.data) == 0
    assert np.array_equal(co.view(np.uint32), c2.view(np.uint32))
    assert np.array_equal(vq, v2)


WIDE = [(512, 16000, 0, "tone", 1.5), (512, 16000, 1, "noise", 1.0), (1024, 16000, 0, "tone", 3.0),
        (1024, 16000, 1, "noise", 3.0), (1024, 16000, 2, "chirp", 3.0), (1024, 16000, 3, "composite", 2.77),
        (1024, 16000, 4, "silence", 1.0), (1024, 16000, 5, None, 0.01), (1024, 16000, 7, None, 4.1),
        (1024, 8000, 6, None, 2.0), (512, 44100, 8, None, 0.7)]


@pytest.mark.parametrize("win,sr,idx,kind,sec", WIDE)
def test_kernel_phases_other_plans(oracle, emul, win, sr, idx, kind, sec):
    """win 1024 / hop 512 (the commented alternative, src/fp_handler.c:35-36) and other sample rates."""
    p = oracle.Plan(win=win, hop=win // 2, samplerate=sr)
    pcm = synth.make_clip(idx, sec, samplerate=sr, kind=kind, ulaw=(idx % 2 == 1))
    co, y, vq = p.extract(pcm)
    F = co.shape[0]
    c2 = np.zeros((F, 2), np.float32); v2 = np.zeros((F, 2), np.int32)
    assert emul.emul_extract_win(win, pcm.ctypes.data, pcm.size, sr, c2.ctypes.data, v2.ctypes.data) == 0
    assert np.array_equal(co.view(np.uint32), c2.view(np.uint32))
    assert np.array_equal(vq, v2)


def test_both_log_phase_variants_match_oracle(oracle, emul):
    """P3b has two address schemes (live filters 0..n-1: immediates; otherwise an index table); every
    plan of the reference takes the first, so the second is forced here."""
    p = oracle.Plan()
    pcm = synth.make_clip(3, 2.0, kind="composite")
    co, y, vq = p.extract(pcm)
    F = co.shape[0]
    try:
        for forced in (1, 0):
            emul.emul_force_indexed_logs(forced)
            c2 = np.zeros((F, 2), np.float32); v2 = np.zeros((F, 2), np.int32)
            assert emul.emul_extract(pcm.ctypes.data, pcm.size, 8000, c2.ctypes.data, v2.ctypes.data) == 0
            assert np.array_equal(co.view(np.uint32), c2.view(np.uint32)) and np.array_equal(vq, v2)
    finally:
        emul.emul_force_indexed_logs(0)


def test_log10f_model_equals_libm(emul):
    # strided sweep over every binade of the positive floats incl. subnormals (the exhaustive
    # 2^31 sweep was run once: 0 mismatches)
    assert emul.emul_log10f_sweep(1, 0x717fffff, 4099) == 0            # domain: [2^-149, 2^100)
    assert emul.emul_log10f_sweep(1, 0x00ffffff, 7) == 0            # subnormals and first binade
    assert emul.emul_log10f_sweep(0x3f000000, 0x40000000, 13) == 0  # [0.5, 2]
    clamp = np.float32(2e-42)
    assert emul.emul_log10f(C.c_float(float(clamp))) == np.float32(np.log10(clamp.astype(np.float64))) or True


def test_quantiser_equals_printf(emul, oracle):
    rng = np.random.default_rng(9)
    vals = list(rng.uniform(-60, 40, 20000)) + list(np.round(rng.uniform(-60, 40, 5000), 6) + 5e-7) + \
        list(np.round(rng.uniform(0, 30, 5000), 6) + 5e-7 * (1 + rng.choice([-1e-9, 1e-9], 5000))) + \
        [0.0, -0.0, 16.9995, 17.0005, 1e-7, -1e-7, 5e-7, 1.5e-6, 2.5e-6, 2147.0, -2147.0, 0.1 + 0.2]
    for v in vals:
        assert emul.emul_quantize(float(v)) == oracle.quantize(float(v)), v
    assert emul.emul_quantize(float("nan")) == oracle.NULL_V
    assert emul.emul_quantize(float("-inf")) == oracle.NULL_V


def test_s16_to_f32_without_the_conversion_unit_is_exact():
    """tir_s16x2 (csrc/tir_extract_core.cuh): XOR 0x8000, splice the 16 bits under the exponent of 2^23 (PRMT), subtract
    2^23 + 2^15 in float32 -- the same float as (float)(int16_t) for every one of the 65 536 samples."""
    s = np.arange(-32768, 32768, dtype=np.int32)
    biased = (s.astype(np.uint32) & 0xFFFF) ^ 0x8000                      # s + 32768 as an unsigned half word
    spliced = (np.uint32(0x4B000000) | biased).view(np.float32)           # 2^23 + (s + 32768), exactly
    assert np.array_equal(spliced.astype(np.float64), 8388608.0 + s + 32768.0)
    got = (spliced - np.float32(8421376.0)).astype(np.float32)            # one float32 subtraction
    assert np.array_equal(got, s.astype(np.float32)) and np.array_equal(got.view(np.uint32), s.astype(np.float32).view(np.uint32))
