"""Seeded synthetic PCM16 corpus: tone / noise / chirp / composite clips,
optional G.711 mu-law round trip ("telephony ulaw-decoded") and runs of exact silence.

Pure numpy, used by tests/ and bench.py to make inputs; nothing here is on the product path.
"""
from __future__ import annotations

import numpy as np

SEED0 = 20180610  # fp_handler.c "Created on: Jun 10, 2018"


def _ulaw_tables():
    # G.711 mu-law, the classic 14-bit biased companding
    BIAS, CLIP = 0x84, 32635
    def enc(x):
        x = x.astype(np.int32)
        sign = (x < 0)
        mag = np.minimum(np.abs(x), CLIP) + BIAS
        exp = np.floor(np.log2(mag)).astype(np.int32) - 7
        exp = np.clip(exp, 0, 7)
        mant = (mag >> (exp + 3)) & 0x0F
        u = ~(np.where(sign, 0x80, 0) | (exp << 4) | mant) & 0xFF
        return u.astype(np.uint8)
    u = np.arange(256, dtype=np.int32)
    v = ~u & 0xFF
    sign, exp, mant = v & 0x80, (v >> 4) & 7, v & 0x0F
    mag = ((mant << 3) + BIAS) << exp
    dec = np.where(sign != 0, BIAS - mag, mag - BIAS).astype(np.int16)
    return enc, dec


_ULAW_ENC, ULAW_DECODE_TABLE = _ulaw_tables()


def ulaw_roundtrip(pcm: np.ndarray) -> np.ndarray:
    return ULAW_DECODE_TABLE[_ULAW_ENC(pcm)]


def ulaw_encode(pcm: np.ndarray) -> np.ndarray:
    return _ULAW_ENC(pcm)


def make_clip(index: int, seconds: float = 30.0, samplerate: int = 8000, kind: str | None = None,
              ulaw: bool = False, silence: bool | None = None) -> np.ndarray:
    """One mono PCM16 clip; seed = SEED0 + index."""
    rng = np.random.default_rng(SEED0 + int(index))
    n = int(round(seconds * samplerate))
    t = np.arange(n, dtype=np.float64) / samplerate
    nyq = samplerate / 2
    if kind is None:
        kind = rng.choice(["tone", "noise", "chirp", "composite"], p=[0.4, 0.3, 0.2, 0.1])
    fmax = 3400.0 if samplerate <= 8000 else 7000.0

    def tone():
        return rng.uniform(0.05, 0.9) * np.sin(2 * np.pi * rng.uniform(200.0, fmax) * t + rng.uniform(0, 2 * np.pi))

    def noise():
        return np.clip(rng.normal(0.0, rng.uniform(0.01, 0.3), n), -1.0, 1.0)

    def chirp():
        f0, f1 = 200.0, 0.85 * nyq
        dur = max(t[-1], 1e-9) if n else 1.0
        return rng.uniform(0.1, 0.8) * np.sin(2 * np.pi * (f0 * t + 0.5 * (f1 - f0) / dur * t * t))

    if kind == "tone":
        x = tone()
    elif kind == "noise":
        x = noise()
    elif kind == "chirp":
        x = chirp()
    elif kind == "composite":
        x = 0.5 * tone() + 0.3 * chirp() + 0.2 * noise()
    elif kind == "silence":
        x = np.zeros(n)
    else:
        raise ValueError(kind)
    pcm = np.round(np.clip(x, -1.0, 1.0) * 32767.0).astype(np.int16)
    if silence is None:
        silence = rng.random() < 0.02
    if silence and n > 8:
        a = int(rng.integers(0, n // 2))
        b = a + int(rng.integers(n // 8, n // 2))
        pcm[a:b] = 0
    if ulaw:
        pcm = ulaw_roundtrip(pcm)
    return pcm


def make_corpus(n_clips: int, seconds: float = 30.0, samplerate: int = 8000, ulaw: bool = False,
                first_index: int = 0, ragged: bool = False):
    """-> (pcm int16 [total], clip_off uint64 [n_clips+1])."""
    clips = []
    for c in range(n_clips):
        s = seconds
        if ragged:
            s = seconds * (0.25 + 0.75 * ((c * 2654435761) % 1000) / 999.0)
        clips.append(make_clip(first_index + c, s, samplerate, ulaw=ulaw))
    off = np.zeros(n_clips + 1, np.uint64)
    off[1:] = np.cumsum([c.size for c in clips])
    pcm = np.concatenate(clips) if clips else np.zeros(0, np.int16)
    return pcm, off


def uuid_for(index: int) -> str:
    """Deterministic RFC-4122 looking lowercase uuid text for audio #index."""
    rng = np.random.default_rng(SEED0 * 7 + int(index))
    b = bytearray(rng.integers(0, 256, 16, dtype=np.uint8).tobytes())
    b[6] = (b[6] & 0x0F) | 0x40
    b[8] = (b[8] & 0x3F) | 0x80
    h = bytes(b).hex()
    return f"{h[0:8]}-{h[8:12]}-{h[12:16]}-{h[16:20]}-{h[20:32]}"
