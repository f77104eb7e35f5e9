"""ctypes binding of the CPU ORACLE (oracle/libtir_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtir_oracle.so")
NULL_V = -(2**31)


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


class _Hit(C.Structure):
    _fields_ = [("found", C.c_int), ("uuid", C.c_char * 64), ("match_count", C.c_int),
                ("frame_count", C.c_int), ("rows_in_windows", C.c_long)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.tiro_plan_create.restype = C.c_void_p
        L.tiro_plan_create.argtypes = [C.c_int] * 5
        L.tiro_plan_destroy.argtypes = [C.c_void_p]
        for name in ("window", "filters", "dct", "band_edges"):
            f = getattr(L, "tiro_plan_" + name)
            f.restype = C.POINTER(C.c_float)
            f.argtypes = [C.c_void_p]
        L.tiro_plan_spec_len.argtypes = [C.c_void_p]
        L.tiro_n_frames.restype = C.c_size_t
        L.tiro_n_frames.argtypes = [C.c_size_t, C.c_int]
        L.tiro_pvoc_norm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.tiro_rfft.argtypes = [C.c_void_p] * 4
        L.tiro_mfcc.argtypes = [C.c_void_p] * 4
        L.tiro_quantize.restype = C.c_int32
        L.tiro_quantize.argtypes = [C.c_double]
        L.tiro_extract.restype = C.c_size_t
        L.tiro_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        L.tiro_extract_interleaved.restype = C.c_size_t
        L.tiro_extract_interleaved.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.tiro_extract_batch.restype = C.c_size_t
        L.tiro_extract_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_int]
        L.tiro_db_open.restype = C.c_void_p
        L.tiro_db_close.argtypes = [C.c_void_p]
        L.tiro_db_sqlite_version.restype = C.c_char_p
        L.tiro_db_add_audio.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p]
        L.tiro_db_add_fingerprints.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_size_t, C.c_int]
        L.tiro_db_delete_audio.argtypes = [C.c_void_p, C.c_char_p]
        L.tiro_db_handle.restype = C.c_void_p
        L.tiro_db_handle.argtypes = [C.c_void_p]
        L.tiro_db_dump_audio.restype = C.c_long
        L.tiro_db_dump_audio.argtypes = [C.c_void_p, C.c_char_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_char_p, C.c_size_t]
        L.tiro_db_count_rows.restype = C.c_long
        L.tiro_db_count_rows.argtypes = [C.c_void_p]
        L.tiro_db_search.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_double,
                                     C.c_int, C.c_int, C.POINTER(_Hit)]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Plan:
    """new_aubio_pvoc(win, hop) + new_aubio_mfcc(win, 40, 2, samplerate) (fp_handler.c:613-617)."""

    def __init__(self, win=512, hop=256, n_filters=40, n_coefs=2, samplerate=8000):
        self.win, self.hop, self.n_filters, self.n_coefs, self.samplerate = win, hop, n_filters, n_coefs, samplerate
        self._h = lib().tiro_plan_create(win, hop, n_filters, n_coefs, samplerate)
        if not self._h:
            raise ValueError("unsupported oracle plan")
        self.L = lib().tiro_plan_spec_len(self._h)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.tiro_plan_destroy(self._h)
            self._h = None

    def _arr(self, name, shape):
        ptr = getattr(lib(), "tiro_plan_" + name)(self._h)
        return np.ctypeslib.as_array(ptr, shape=shape).copy()

    @property
    def window(self):
        return self._arr("window", (self.win,))

    @property
    def filters(self):
        return self._arr("filters", (self.n_filters, self.L))

    @property
    def dct(self):
        return self._arr("dct", (self.n_coefs, self.n_filters))

    @property
    def band_edges(self):
        return self._arr("band_edges", (self.n_filters + 2,))

    def n_frames(self, n_samples):
        return int(lib().tiro_n_frames(int(n_samples), self.hop))

    def pvoc_norm(self, data):
        data = np.ascontiguousarray(data, dtype=np.float32)
        assert data.shape == (self.win,)
        out = np.empty(self.L, np.float32)
        lib().tiro_pvoc_norm(self._h, _p(data), _p(out))
        return out

    def rfft(self, frame):
        frame = np.ascontiguousarray(frame, dtype=np.float32)
        assert frame.shape == (self.win,)
        re = np.empty(self.L, np.float32)
        im = np.empty(self.L, np.float32)
        lib().tiro_rfft(self._h, _p(frame), _p(re), _p(im))
        return re, im

    def mfcc(self, norm):
        norm = np.ascontiguousarray(norm, dtype=np.float32)
        mel = np.empty(self.n_filters, np.float32)
        coef = np.empty(self.n_coefs, np.float32)
        lib().tiro_mfcc(self._h, _p(norm), _p(mel), _p(coef))
        return mel, coef

    def extract(self, pcm):
        """-> coef[F, n_coefs] f32, y[F, n_coefs] f64, vq[F, n_coefs] i32 (NULL_V = NULL column)."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        F = self.n_frames(pcm.size)
        coef = np.empty((F, self.n_coefs), np.float32)
        y = np.empty((F, self.n_coefs), np.float64)
        vq = np.empty((F, self.n_coefs), np.int32)
        lib().tiro_extract(self._h, _p(pcm), pcm.size, _p(coef), _p(y), _p(vq))
        return coef, y, vq

    def extract_interleaved(self, pcm, channels):
        """pcm[sample frame, channel] (or flat, interleaved) -> the same triple, from the channels' float mean."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1)
        n = pcm.size // channels
        F = self.n_frames(n)
        coef = np.empty((F, self.n_coefs), np.float32)
        y = np.empty((F, self.n_coefs), np.float64)
        vq = np.empty((F, self.n_coefs), np.int32)
        lib().tiro_extract_interleaved(self._h, _p(pcm), n, channels, _p(coef), _p(y), _p(vq))
        return coef, y, vq

    def extract_batch(self, pcm, clip_off, n_threads=1, want_y=True):
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        clip_off = np.ascontiguousarray(clip_off, dtype=np.uint64)
        n_clips = clip_off.size - 1
        lens = np.diff(clip_off.astype(np.int64))
        F = int(((lens + self.hop - 1) // self.hop).sum())
        coef = np.empty((F, self.n_coefs), np.float32)
        y = np.empty((F, self.n_coefs), np.float64) if want_y else None
        vq = np.empty((F, self.n_coefs), np.int32)
        got = lib().tiro_extract_batch(self._h, _p(pcm), _p(clip_off), n_clips, _p(coef), _p(y), _p(vq), n_threads)
        assert got == F
        return coef, y, vq


class LibAubio:
    """The REAL libaubio, if one is ever installed (it is not in this image: SURVEY.md 8c): the same calls
    src/fp_handler.c:604-668 makes -- new_aubio_pvoc / new_aubio_mfcc / aubio_pvoc_do / aubio_mfcc_do -- on hops
    fed the way aubio_source does (PCM16 / 32768, last hop zero padded).  tests/test_oracle_extract.py pins the
    restatement against it the day the library exists; until then parity at the libaubio boundary is unpinned."""

    class _FVec(C.Structure):
        _fields_ = [("length", C.c_uint), ("data", C.POINTER(C.c_float))]

    class _CVec(C.Structure):
        _fields_ = [("length", C.c_uint), ("norm", C.POINTER(C.c_float)), ("phas", C.POINTER(C.c_float))]

    _lib = None

    @classmethod
    def load(cls):
        """-> the CDLL or None"""
        if cls._lib is None:
            for name in (os.environ.get("TIR_LIBAUBIO"), "libaubio.so.5", "libaubio.so"):
                if not name:
                    continue
                try:
                    L = C.CDLL(name)
                except OSError:
                    continue
                L.new_fvec.restype = C.POINTER(cls._FVec); L.new_fvec.argtypes = [C.c_uint]
                L.new_cvec.restype = C.POINTER(cls._CVec); L.new_cvec.argtypes = [C.c_uint]
                L.new_aubio_pvoc.restype = C.c_void_p; L.new_aubio_pvoc.argtypes = [C.c_uint, C.c_uint]
                L.new_aubio_mfcc.restype = C.c_void_p; L.new_aubio_mfcc.argtypes = [C.c_uint] * 4
                L.aubio_pvoc_do.argtypes = [C.c_void_p, C.POINTER(cls._FVec), C.POINTER(cls._CVec)]
                L.aubio_mfcc_do.argtypes = [C.c_void_p, C.POINTER(cls._CVec), C.POINTER(cls._FVec)]
                for f, t in (("del_aubio_pvoc", C.c_void_p), ("del_aubio_mfcc", C.c_void_p), ("del_fvec", C.POINTER(cls._FVec)), ("del_cvec", C.POINTER(cls._CVec))):
                    getattr(L, f).argtypes = [t]
                cls._lib = L
                break
        return cls._lib

    @classmethod
    def available(cls) -> bool:
        return cls.load() is not None

    def __init__(self, win=512, hop=256, n_filters=40, n_coefs=2, samplerate=8000):
        self.L = self.load()
        if self.L is None:
            raise RuntimeError("libaubio is not installed")
        self.win, self.hop, self.n_filters, self.n_coefs, self.sr = win, hop, n_filters, n_coefs, samplerate

    def extract(self, pcm):
        """create_audio_fingerprints() (src/fp_handler.c:577-671) for in-memory PCM16 -> coef f32 [F,2], y f64 [F,2]"""
        L = self.L
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        F = (pcm.size + self.hop - 1) // self.hop
        pv, mf = L.new_aubio_pvoc(self.win, self.hop), L.new_aubio_mfcc(self.win, self.n_filters, self.n_coefs, self.sr)
        buf, out, grain = L.new_fvec(self.hop), L.new_fvec(self.n_coefs), L.new_cvec(self.win)
        coef = np.zeros((F, self.n_coefs), np.float32)
        x = np.zeros(F * self.hop, np.float32)
        x[:pcm.size] = pcm.astype(np.float32) / np.float32(32768.0)
        try:
            for f in range(F):
                C.memmove(buf.contents.data, x[f * self.hop:].ctypes.data, self.hop * 4)
                L.aubio_pvoc_do(pv, buf, grain)
                L.aubio_mfcc_do(mf, grain, out)
                coef[f] = np.ctypeslib.as_array(out.contents.data, (self.n_coefs,))
        finally:
            L.del_aubio_pvoc(pv), L.del_aubio_mfcc(mf), L.del_fvec(buf), L.del_fvec(out), L.del_cvec(grain)
        with np.errstate(divide="ignore"):
            y = 10.0 * np.log10(np.abs(coef.astype(np.float64)))
        return coef, y


def set_fft_kind(kind: int) -> None:
    """0 = TIR-FFT (default), 1 = Ooura-style, 2 = float64 rounded (FFT-order sensitivity study)"""
    lib().tiro_set_fft_kind(int(kind))


def quantize(y: float) -> int:
    return int(lib().tiro_quantize(float(y)))


class SqliteDB:
    """The reference's in-memory SQLite DB, driven with the reference's literal SQL text."""

    def __init__(self):
        self._h = lib().tiro_db_open()
        if not self._h:
            raise RuntimeError("sqlite oracle could not open")

    def close(self):
        if self._h and _lib is not None:
            _lib.tiro_db_close(self._h)
        self._h = None

    __del__ = close

    @staticmethod
    def sqlite_version():
        return lib().tiro_db_sqlite_version().decode()

    def add_audio(self, uuid, y, context="ctx", name=None, hash_="0" * 32, literal_autocommit=False):
        y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1, 2)
        rc = lib().tiro_db_add_audio(self._h, uuid.encode(), (name or uuid + ".wav").encode(), context.encode(),
                                     hash_.encode())
        rc |= lib().tiro_db_add_fingerprints(self._h, context.encode(), uuid.encode(), _p(y), y.shape[0],
                                             1 if literal_autocommit else 0)
        if rc:
            raise RuntimeError("sqlite oracle insert failed")

    def delete_audio(self, uuid):
        if lib().tiro_db_delete_audio(self._h, uuid.encode()):
            raise RuntimeError("sqlite oracle delete failed")

    @property
    def handle(self):
        """the raw sqlite3* (int), as the module would pass g_db_ctx->db"""
        return lib().tiro_db_handle(self._h)

    def dump_audio(self, uuid, cap=100000):
        """-> (frame_idx, max1, max2, type1, type2, context) of the rows of one audio, rowid order"""
        fi = np.zeros(cap, np.int64); m1 = np.zeros(cap, np.float64); m2 = np.zeros(cap, np.float64)
        t1 = np.zeros(cap, np.int32); t2 = np.zeros(cap, np.int32)
        ctxbuf = C.create_string_buffer(512)
        n = lib().tiro_db_dump_audio(self._h, uuid.encode(), cap, _p(fi), _p(m1), _p(m2), _p(t1), _p(t2), ctxbuf, 512)
        assert 0 <= n <= cap
        return fi[:n], m1[:n], m2[:n], t1[:n], t2[:n], ctxbuf.value.decode()

    def count_rows(self):
        return int(lib().tiro_db_count_rows(self._h))

    def search(self, y, coefs=1, tolerance=0.001, freq_ignore_low=-1, freq_ignore_high=-1, has_y=None):
        """fp_search_fingerprint_info on extracted query values -> dict or None (NOTFOUND)."""
        y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1, 2)
        hy = None if has_y is None else np.ascontiguousarray(has_y, dtype=np.uint8).reshape(-1, 2)
        hit = _Hit()
        rc = lib().tiro_db_search(self._h, _p(y), _p(hy), y.shape[0], coefs, float(tolerance),
                                  int(freq_ignore_low), int(freq_ignore_high), C.byref(hit))
        if rc:
            return None  # argument rejected (coefs out of range): reference returns NULL
        if not hit.found:
            return None
        return {"uuid": hit.uuid.decode(), "match_count": hit.match_count, "frame_count": hit.frame_count,
                "votes_total": hit.rows_in_windows}
