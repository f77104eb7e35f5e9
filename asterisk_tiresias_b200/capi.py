"""ctypes binding of libtiresias_gpu.so (include/tiresias_gpu.h) for tests/ and bench.py.

No compute happens here: every call goes through the C ABI.  If the library is missing or there
is no sm_100 GPU the calls fail loudly -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import uuid as _uuid

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtiresias_gpu.so")
NULL_V = -(2**31)

OK, ERR_ARG, ERR_CUDA, ERR_NOMEM, ERR_STATE, ERR_NOTFOUND = 0, -1, -2, -3, -4, -5


class TirError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"tiresias_gpu error {code}: {msg}")
        self.code = code


class Cfg(C.Structure):
    _fields_ = [("device", C.c_int), ("win", C.c_int), ("hop", C.c_int), ("n_filters", C.c_int),
                ("samplerate", C.c_int), ("stream", C.c_void_p)]


class Hit(C.Structure):
    _fields_ = [("uuid", C.c_uint8 * 16), ("match_count", C.c_int32), ("frame_count", C.c_int32)]


HIT_DTYPE = np.dtype([("uuid", np.uint8, 16), ("match_count", np.int32), ("frame_count", np.int32)])
assert HIT_DTYPE.itemsize == C.sizeof(Hit) == 24

_lib = None


def lib():
    """Load the shared library (raises if it has not been built: run __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TirError(ERR_STATE, f"{LIB_PATH} is missing: build it with `python -m asterisk_tiresias_b200.build`")
        L = C.CDLL(LIB_PATH)
        vp, u64p = C.c_void_p, C.c_void_p
        L.tir_abi_version.restype = C.c_int
        L.tir_cfg_default.argtypes = [C.POINTER(Cfg)]
        L.tir_open.argtypes = [C.POINTER(Cfg), C.POINTER(vp)]
        L.tir_close.argtypes = [vp]
        L.tir_last_error.restype = C.c_char_p
        L.tir_last_error.argtypes = [vp]
        L.tir_n_frames.restype = C.c_uint64
        L.tir_n_frames.argtypes = [C.c_uint64, C.c_int]
        L.tir_extract.argtypes = [vp, vp, u64p, C.c_uint32, vp, vp, C.POINTER(C.c_uint64)]
        L.tir_selftest.argtypes = [vp, u64p, C.c_uint32, C.c_uint32, C.c_uint32, vp]
        L.tir_extract_ulaw.argtypes = [vp, vp, u64p, C.c_uint32, vp, vp, C.POINTER(C.c_uint64)]
        L.tir_extract_interleaved.argtypes = [vp, vp, C.c_int, u64p, C.c_uint32, vp, vp, C.POINTER(C.c_uint64)]
        L.tir_extract_dev.argtypes = [vp, vp, u64p, C.c_uint32, vp, vp, C.POINTER(C.c_uint64)]
        L.tir_get_tables.argtypes = [vp, vp, vp, vp]
        L.tir_launch_count.restype = C.c_uint64
        L.tir_launch_count.argtypes = [vp]
        if hasattr(L, "tir_db_load"):
            L.tir_db_load.argtypes = [vp, C.c_uint32, vp, u64p, vp, vp]
            L.tir_db_load_dev.argtypes = [vp, C.c_uint32, vp, vp, vp, vp, C.c_uint64]
            L.tir_set_profiling.argtypes = [vp, C.c_int]
            L.tir_last_kernel_ms.restype = C.c_float
            L.tir_last_kernel_ms.argtypes = [vp, C.c_int]
            L.tir_db_add.argtypes = [vp, vp, vp, vp, C.c_uint32]
            L.tir_db_remove.argtypes = [vp, vp]
            L.tir_db_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
            L.tir_match.argtypes = [vp, vp, u64p, C.c_uint32, C.c_int, C.c_double, C.c_int, C.c_int, vp]
            L.tir_match_dev.argtypes = [vp, vp, u64p, C.c_uint32, C.c_int, C.c_double, C.c_int, C.c_int, vp]
            L.tir_search.argtypes = [vp, vp, u64p, C.c_uint32, C.c_int, C.c_double, C.c_int, C.c_int, vp]
            L.tir_merge_hits_dev.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp]
            L.tir_shard_of.restype = C.c_uint32
            L.tir_shard_of.argtypes = [vp, C.c_uint32]
            L.tir_db_load_sqlite.argtypes = [vp, vp, u64p, u64p, u64p]
            L.tir_db_load_sqlite_file.argtypes = [vp, C.c_char_p, u64p, u64p, u64p]
            L.tir_sqlite_insert_fingerprints.argtypes = [vp, vp, C.c_char_p, C.c_char_p, vp, C.c_uint32]
            L.tir_group_open.argtypes = [vp, vp, C.c_int, C.POINTER(vp)]
            L.tir_group_close.argtypes = [vp]
            L.tir_group_last_error.restype = C.c_char_p
            L.tir_group_last_error.argtypes = [vp]
            L.tir_group_size.argtypes = [vp]
            L.tir_group_ctx.restype = vp
            L.tir_group_ctx.argtypes = [vp, C.c_int]
            L.tir_group_db_load.argtypes = [vp, C.c_uint32, vp, u64p, vp, vp]
            L.tir_group_db_add.argtypes = [vp, vp, vp, vp, C.c_uint32]
            L.tir_group_db_remove.argtypes = [vp, vp]
            L.tir_group_db_stats.argtypes = [vp, u64p, u64p]
            L.tir_group_search.argtypes = [vp, vp, u64p, C.c_uint32, C.c_int, C.c_double, C.c_int, C.c_int, vp]
            L.tir_p2p_create.argtypes = [vp, C.c_int, C.c_int, C.c_uint32, C.POINTER(vp)]
            L.tir_p2p_create2.argtypes = [vp, C.c_int, C.c_int, C.c_uint32, C.c_uint64, C.POINTER(vp)]
            L.tir_p2p_search.argtypes = [vp, vp, u64p, C.c_uint32, C.c_uint32, u64p, C.c_uint32, C.c_int, C.c_double, C.c_int, C.c_int, vp, vp]
            L.tir_group_stats.argtypes = [vp, vp, vp]
            L.tir_group_batcher_start.argtypes = [vp, C.c_uint32, C.c_uint32]
            L.tir_group_batcher_stop.argtypes = [vp]
            L.tir_group_search_one.argtypes = [vp, vp, C.c_uint64, C.c_int, C.c_double, C.c_int, C.c_int, vp]
            L.tir_group_batcher_stats.argtypes = [vp, vp, vp, vp]
            L.tir_p2p_reserve.argtypes = [vp, C.c_uint64]
            L.tir_db_index_stats.argtypes = [vp, vp, vp, vp, vp]
            L.tir_match_graph_stats.argtypes = [vp, vp, vp]
            L.tir_p2p_handle.argtypes = [vp, vp]
            L.tir_p2p_connect.argtypes = [vp, vp]
            L.tir_p2p_connect_local.argtypes = [vp, vp]
            L.tir_p2p_match_dev.argtypes = [vp, vp, u64p, C.c_uint32, C.c_int, C.c_double, C.c_int, C.c_int, vp]
            L.tir_p2p_error.argtypes = [vp, vp]
            L.tir_p2p_destroy.argtypes = [vp]
            L.tir_p2p_destroy.restype = None
            L.tir_batcher_start.argtypes = [vp, C.c_uint32, C.c_uint32]
            L.tir_batcher_stop.argtypes = [vp]
            L.tir_search_one.argtypes = [vp, vp, C.c_uint64, C.c_int, C.c_double, C.c_int, C.c_int, vp]
            L.tir_batcher_stats.argtypes = [vp, u64p, u64p, u64p]
            L.tir_stream_open.argtypes = [vp, C.POINTER(vp)]
            L.tir_stream_feed.argtypes = [vp, vp, C.c_uint32]
            L.tir_stream_samples.restype = C.c_uint64
            L.tir_stream_samples.argtypes = [vp]
            L.tir_stream_finish.argtypes = [vp, C.c_int, C.c_double, C.c_int, C.c_int, vp]
            L.tir_stream_close.argtypes = [vp]
            L.tir_stream_frames_done.restype = C.c_uint64
            L.tir_stream_frames_done.argtypes = [vp]
            L.tir_stream_stats.argtypes = [vp, vp, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def uuid_to_bytes(text: str) -> np.ndarray:
    return np.frombuffer(_uuid.UUID(text).bytes, dtype=np.uint8).copy()


def bytes_to_uuid(b) -> str:
    return str(_uuid.UUID(bytes=bytes(bytearray(np.asarray(b, dtype=np.uint8).tolist()))))


class Context:
    """tir_ctx wrapper.  `stream` is a raw cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream)."""

    def __init__(self, device=0, win=512, hop=256, n_filters=40, samplerate=8000, stream=None):
        L = lib()
        cfg = Cfg()
        L.tir_cfg_default(C.byref(cfg))
        cfg.device, cfg.win, cfg.hop, cfg.n_filters, cfg.samplerate = device, win, hop, n_filters, samplerate
        cfg.stream = stream
        self.cfg = cfg
        self._h = C.c_void_p()
        rc = L.tir_open(C.byref(cfg), C.byref(self._h))
        if rc != OK:
            msg = L.tir_last_error(self._h).decode() if self._h else "tir_open failed"
            if self._h:
                L.tir_close(self._h)
            self._h = None
            raise TirError(rc, msg)

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.tir_close(self._h)
        self._h = None

    __del__ = close

    def _chk(self, rc):
        if rc != OK:
            raise TirError(rc, lib().tir_last_error(self._h).decode())

    @property
    def hop(self):
        return self.cfg.hop

    @property
    def launches(self):
        return int(lib().tir_launch_count(self._h))

    def tables(self):
        win, nf = self.cfg.win, self.cfg.n_filters
        w = np.empty(win, np.float32)
        fb = np.empty((nf, win // 2 + 1), np.float32)
        d = np.empty((2, nf), np.float32)
        self._chk(lib().tir_get_tables(self._h, _p(w), _p(fb), _p(d)))
        return w, fb, d

    # ---- extraction -------------------------------------------------------------------------
    def n_frames(self, clip_off):
        lens = np.diff(np.asarray(clip_off, dtype=np.uint64).astype(np.int64))
        return int(((lens + self.hop - 1) // self.hop).sum())

    def extract(self, pcm, clip_off=None):
        """Host buffers in, host buffers out: coef [F,2] f32, vq [F,2] i32."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        if clip_off is None:
            clip_off = np.array([0, pcm.size], np.uint64)
        clip_off = np.ascontiguousarray(clip_off, dtype=np.uint64)
        F = self.n_frames(clip_off)
        coef = np.empty((F, 2), np.float32)
        vq = np.empty((F, 2), np.int32)
        nf = C.c_uint64()
        self._chk(lib().tir_extract(self._h, _p(pcm), _p(clip_off), clip_off.size - 1, _p(coef), _p(vq), C.byref(nf)))
        assert nf.value == F
        return coef, vq

    def selftest_sqrt(self):
        bad = C.c_uint64(1 << 60)
        self._chk(lib().tir_selftest(self._h, C.byref(bad), 0, 0, 0, None))
        return int(bad.value)

    def selftest_log10f(self, first_bits, step, count):
        out = np.empty(count, np.float32)
        self._chk(lib().tir_selftest(self._h, None, first_bits, step, count, _p(out)))
        return out

    def extract_ulaw(self, ulaw, clip_off=None):
        """G.711 mu-law bytes in (uint8), host buffers out."""
        ulaw = np.ascontiguousarray(ulaw, dtype=np.uint8)
        if clip_off is None:
            clip_off = np.array([0, ulaw.size], np.uint64)
        clip_off = np.ascontiguousarray(clip_off, dtype=np.uint64)
        F = self.n_frames(clip_off)
        coef = np.empty((F, 2), np.float32)
        vq = np.empty((F, 2), np.int32)
        nf = C.c_uint64()
        self._chk(lib().tir_extract_ulaw(self._h, _p(ulaw), _p(clip_off), clip_off.size - 1, _p(coef), _p(vq), C.byref(nf)))
        assert nf.value == F
        return coef, vq

    def extract_interleaved(self, pcm, channels, clip_off=None):
        """pcm[sample frame, channel] int16 (interleaved); clip_off in sample frames."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1)
        if clip_off is None:
            clip_off = np.array([0, pcm.size // channels], np.uint64)
        clip_off = np.ascontiguousarray(clip_off, dtype=np.uint64)
        F = self.n_frames(clip_off)
        coef = np.empty((F, 2), np.float32)
        vq = np.empty((F, 2), np.int32)
        nf = C.c_uint64()
        self._chk(lib().tir_extract_interleaved(self._h, _p(pcm), channels, _p(clip_off), clip_off.size - 1, _p(coef), _p(vq), C.byref(nf)))
        assert nf.value == F
        return coef, vq

    def extract_dev(self, d_pcm_ptr, clip_off, d_coef_ptr, d_vq_ptr):
        """Device pointers (ints); asynchronous on the context's stream.  Returns frame count."""
        clip_off = np.ascontiguousarray(clip_off, dtype=np.uint64)
        nf = C.c_uint64()
        self._chk(lib().tir_extract_dev(self._h, C.c_void_p(d_pcm_ptr), _p(clip_off), clip_off.size - 1,
                                        C.c_void_p(d_coef_ptr), C.c_void_p(d_vq_ptr), C.byref(nf)))
        return int(nf.value)

    # ---- device DB --------------------------------------------------------------------------
    def db_load(self, uuids, row_off, v1, v2):
        """uuids: [n,16] uint8; row_off [n+1] uint64; v1/v2 int32 micro-units (NULL_V = NULL)."""
        uuids = np.ascontiguousarray(uuids, dtype=np.uint8).reshape(-1, 16)
        row_off = np.ascontiguousarray(row_off, dtype=np.uint64)
        v1 = np.ascontiguousarray(v1, dtype=np.int32)
        v2 = np.ascontiguousarray(v2, dtype=np.int32)
        assert row_off.size == uuids.shape[0] + 1 and v1.size == v2.size == int(row_off[-1])
        self._chk(lib().tir_db_load(self._h, uuids.shape[0], _p(uuids), _p(row_off), _p(v1), _p(v2)))

    def db_load_dev(self, n_audio, d_uuid_ptr, d_row_off_ptr, d_v1_ptr, d_v2_ptr, n_rows):
        self._chk(lib().tir_db_load_dev(self._h, n_audio, C.c_void_p(d_uuid_ptr), C.c_void_p(d_row_off_ptr),
                                        C.c_void_p(d_v1_ptr), C.c_void_p(d_v2_ptr), n_rows))

    def set_profiling(self, on=True):
        lib().tir_set_profiling(self._h, 1 if on else 0)

    def last_kernel_ms(self, which=0):
        return float(lib().tir_last_kernel_ms(self._h, which))

    def db_add(self, uuid16, v1, v2):
        u = np.ascontiguousarray(uuid16, dtype=np.uint8)
        v1 = np.ascontiguousarray(v1, dtype=np.int32)
        v2 = np.ascontiguousarray(v2, dtype=np.int32)
        self._chk(lib().tir_db_add(self._h, _p(u), _p(v1), _p(v2), v1.size))

    def db_remove(self, uuid16):
        u = np.ascontiguousarray(uuid16, dtype=np.uint8)
        self._chk(lib().tir_db_remove(self._h, _p(u)))

    def db_stats(self):
        a, r = C.c_uint64(), C.c_uint64()
        self._chk(lib().tir_db_stats(self._h, C.byref(a), C.byref(r)))
        return int(a.value), int(r.value)

    def db_index_stats(self):
        """-> dict(full_builds, tail_builds, tail_audios, tombstones)"""
        v = [C.c_uint64() for _ in range(4)]
        self._chk(lib().tir_db_index_stats(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("full_builds", "tail_builds", "tail_audios", "tombstones"), (int(x.value) for x in v)))

    def match_graph_stats(self):
        """-> dict(graph_launches, graphs_built): batches replayed as one CUDA graph launch, graphs captured"""
        a, b = C.c_uint64(), C.c_uint64()
        self._chk(lib().tir_match_graph_stats(self._h, C.byref(a), C.byref(b)))
        return {"graph_launches": int(a.value), "graphs_built": int(b.value)}

    # ---- match ------------------------------------------------------------------------------
    def match(self, y, frame_off=None, coefs=1, tolerance=0.001, freq_ignore_low=-1, freq_ignore_high=-1):
        y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1, 2)
        if frame_off is None:
            frame_off = np.array([0, y.shape[0]], np.uint64)
        frame_off = np.ascontiguousarray(frame_off, dtype=np.uint64)
        nq = frame_off.size - 1
        hits = np.zeros(nq, HIT_DTYPE)
        self._chk(lib().tir_match(self._h, _p(y), _p(frame_off), nq, coefs, float(tolerance), int(freq_ignore_low),
                                  int(freq_ignore_high), _p(hits)))
        return hits

    def match_dev(self, d_coef_ptr, frame_off, d_hits_ptr, coefs=1, tolerance=0.001, freq_ignore_low=-1,
                  freq_ignore_high=-1):
        frame_off = np.ascontiguousarray(frame_off, dtype=np.uint64)
        self._chk(lib().tir_match_dev(self._h, C.c_void_p(d_coef_ptr), _p(frame_off), frame_off.size - 1, coefs,
                                      float(tolerance), int(freq_ignore_low), int(freq_ignore_high),
                                      C.c_void_p(d_hits_ptr)))

    def search(self, pcm, clip_off=None, coefs=1, tolerance=0.001, freq_ignore_low=-1, freq_ignore_high=-1):
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        if clip_off is None:
            clip_off = np.array([0, pcm.size], np.uint64)
        clip_off = np.ascontiguousarray(clip_off, dtype=np.uint64)
        nq = clip_off.size - 1
        hits = np.zeros(nq, HIT_DTYPE)
        self._chk(lib().tir_search(self._h, _p(pcm), _p(clip_off), nq, coefs, float(tolerance), int(freq_ignore_low),
                                   int(freq_ignore_high), _p(hits)))
        return hits

    # ---- SQLite mirror ------------------------------------------------------------------------
    def db_load_sqlite(self, sqlite3_handle=None, path=None):
        """-> (n_audio, n_rows, n_skipped)"""
        a, r, k = C.c_uint64(), C.c_uint64(), C.c_uint64()
        if path is not None:
            self._chk(lib().tir_db_load_sqlite_file(self._h, path.encode(), C.byref(a), C.byref(r), C.byref(k)))
        else:
            self._chk(lib().tir_db_load_sqlite(self._h, C.c_void_p(sqlite3_handle), C.byref(a), C.byref(r), C.byref(k)))
        return int(a.value), int(r.value), int(k.value)

    # ---- concurrent front-end ---------------------------------------------------------------
    def batcher_start(self, max_batch=1024, max_wait_us=200):
        self._chk(lib().tir_batcher_start(self._h, max_batch, max_wait_us))

    def batcher_stop(self):
        self._chk(lib().tir_batcher_stop(self._h))

    def batcher_stats(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._chk(lib().tir_batcher_stats(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return int(a.value), int(b.value), int(c.value)

    def search_one(self, pcm, coefs=1, tolerance=0.001, freq_ignore_low=-1, freq_ignore_high=-1):
        """Blocking single-recording search (ctypes releases the GIL: callable from many threads)."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        hit = np.zeros(1, HIT_DTYPE)
        self._chk(lib().tir_search_one(self._h, _p(pcm), pcm.size, coefs, float(tolerance), int(freq_ignore_low),
                                       int(freq_ignore_high), _p(hit)))
        return hit[0]

    def stream(self):
        return Stream(self)

    def merge_hits_dev(self, d_gathered_ptr, n_shards, n_queries, d_out_ptr):
        self._chk(lib().tir_merge_hits_dev(self._h, C.c_void_p(d_gathered_ptr), n_shards, n_queries,
                                           C.c_void_p(d_out_ptr)))


def shard_of(uuid16, n_shards: int) -> int:
    u = np.ascontiguousarray(uuid16, dtype=np.uint8)
    return int(lib().tir_shard_of(_p(u), n_shards))


class P2P:
    """tir_p2p wrapper: the cross-GPU winner exchange of the sharded match over NVLink peer memory."""

    def __init__(self, ctx, rank, world, max_queries, max_frames=0):
        """max_frames > 0 also reserves the coefficient buffers of the sharded search (search())."""
        self.ctx, self.rank, self.world = ctx, rank, world
        self._p = C.c_void_p()
        ctx._chk(lib().tir_p2p_create2(ctx._h, rank, world, max_queries, int(max_frames), C.byref(self._p)))

    def handle(self) -> bytes:
        buf = (C.c_ubyte * 64)()
        self.ctx._chk(lib().tir_p2p_handle(self._p, buf))
        return bytes(buf)

    def connect(self, handles):
        """handles: the 64-byte handles of all ranks, in rank order (one process per GPU)."""
        blob = b"".join(handles)
        assert len(blob) == 64 * self.world
        arr = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self.ctx._chk(lib().tir_p2p_connect(self._p, arr))

    def connect_local(self, all_p2p):
        """all_p2p: the P2P objects of all ranks of THIS process, in rank order."""
        arr = (C.c_void_p * self.world)(*[x._p for x in all_p2p])
        self.ctx._chk(lib().tir_p2p_connect_local(self._p, arr))

    def match_dev(self, d_coef, frame_off, d_final, coefs=1, tolerance=0.001, low=-1, high=-1):
        foff = np.ascontiguousarray(frame_off, dtype=np.uint64)
        self.ctx._chk(lib().tir_p2p_match_dev(self._p, C.c_void_p(d_coef), _p(foff), foff.size - 1, coefs, float(tolerance), int(low), int(high),
                                              C.c_void_p(d_final)))

    def reserve(self, max_local_samples):
        """pre-size the context's scratch (needed when one host thread drives several ranks)"""
        self.ctx._chk(lib().tir_p2p_reserve(self._p, int(max_local_samples)))

    def search(self, pcm, clip_off, first_query, all_frame_off, coefs=1, tolerance=0.001, low=-1, high=-1, d_final=None,
               want_hits=True, pcm_ptr=None, hits_ptr=None):
        """tir_p2p_search: this rank's slice of the batch's clips (host PCM16; or pcm_ptr = address of a pinned
        buffer) -> the winners of ALL queries.  Returns the hits array (or None with want_hits=False / hits_ptr)."""
        off = np.ascontiguousarray(clip_off, dtype=np.uint64)
        foff = np.ascontiguousarray(all_frame_off, dtype=np.uint64)
        n_local, n_total = max(off.size - 1, 0), foff.size - 1
        if pcm_ptr is None:
            pcm = np.ascontiguousarray(pcm, dtype=np.int16)
            pcm_ptr = pcm.ctypes.data
        hits = None
        if hits_ptr is None and want_hits:
            hits = np.zeros(n_total, dtype=HIT_DTYPE)
            hits_ptr = hits.ctypes.data
        self.ctx._chk(lib().tir_p2p_search(self._p, C.c_void_p(pcm_ptr), _p(off), n_local, int(first_query), _p(foff), n_total, coefs,
                                           float(tolerance), int(low), int(high), C.c_void_p(hits_ptr or 0), C.c_void_p(d_final or 0)))
        return hits

    def error(self) -> int:
        e = C.c_uint32(0)
        self.ctx._chk(lib().tir_p2p_error(self._p, C.byref(e)))
        return int(e.value)

    def close(self):
        if getattr(self, "_p", None) and _lib is not None:
            _lib.tir_p2p_destroy(self._p)
        self._p = None

    __del__ = close


class Group:
    """tir_group wrapper: one context per device in this process, table sharded by uuid."""

    def __init__(self, devices, win=512, hop=256, n_filters=40, samplerate=8000):
        L = lib()
        cfg = Cfg()
        L.tir_cfg_default(C.byref(cfg))
        cfg.win, cfg.hop, cfg.n_filters, cfg.samplerate = win, hop, n_filters, samplerate
        dev = np.ascontiguousarray(devices, dtype=np.int32)
        self._g = C.c_void_p()
        rc = L.tir_group_open(C.byref(cfg), _p(dev), dev.size, C.byref(self._g))
        if rc != OK:
            msg = L.tir_group_last_error(self._g).decode() if self._g else "tir_group_open failed"
            if self._g:
                L.tir_group_close(self._g)
            self._g = None
            raise TirError(rc, msg)

    def close(self):
        if getattr(self, "_g", None) and _lib is not None:
            _lib.tir_group_close(self._g)
        self._g = None

    __del__ = close

    def _chk(self, rc):
        if rc != OK:
            raise TirError(rc, lib().tir_group_last_error(self._g).decode())

    def db_load(self, uuids, row_off, v1, v2):
        uuids = np.ascontiguousarray(uuids, dtype=np.uint8).reshape(-1, 16)
        row_off = np.ascontiguousarray(row_off, dtype=np.uint64)
        v1 = np.ascontiguousarray(v1, dtype=np.int32); v2 = np.ascontiguousarray(v2, dtype=np.int32)
        self._chk(lib().tir_group_db_load(self._g, uuids.shape[0], _p(uuids), _p(row_off), _p(v1), _p(v2)))

    def db_add(self, uuid, v1, v2):
        uuid = np.ascontiguousarray(uuid, dtype=np.uint8)
        v1 = np.ascontiguousarray(v1, dtype=np.int32); v2 = np.ascontiguousarray(v2, dtype=np.int32)
        self._chk(lib().tir_group_db_add(self._g, _p(uuid), _p(v1), _p(v2), v1.size))

    def db_remove(self, uuid):
        uuid = np.ascontiguousarray(uuid, dtype=np.uint8)
        self._chk(lib().tir_group_db_remove(self._g, _p(uuid)))

    def db_stats(self):
        a, r = C.c_uint64(), C.c_uint64()
        self._chk(lib().tir_group_db_stats(self._g, C.byref(a), C.byref(r)))
        return int(a.value), int(r.value)

    def stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        self._chk(lib().tir_group_stats(self._g, C.byref(a), C.byref(b)))
        return {"fused": int(a.value), "copy_path": int(b.value)}

    def batcher_start(self, max_batch=1024, max_wait_us=300):
        self._chk(lib().tir_group_batcher_start(self._g, max_batch, max_wait_us))

    def batcher_stop(self):
        self._chk(lib().tir_group_batcher_stop(self._g))

    def batcher_stats(self):
        v = [C.c_uint64() for _ in range(3)]
        self._chk(lib().tir_group_batcher_stats(self._g, *[C.byref(x) for x in v]))
        return tuple(int(x.value) for x in v)

    def search_one(self, pcm, coefs=1, tolerance=0.001, freq_ignore_low=-1, freq_ignore_high=-1):
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        hit = np.zeros(1, HIT_DTYPE)
        self._chk(lib().tir_group_search_one(self._g, _p(pcm), pcm.size, coefs, float(tolerance), int(freq_ignore_low), int(freq_ignore_high), _p(hit)))
        return hit[0]

    def search(self, pcm, clip_off=None, coefs=1, tolerance=0.001, freq_ignore_low=-1, freq_ignore_high=-1):
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        if clip_off is None:
            clip_off = np.array([0, pcm.size], np.uint64)
        clip_off = np.ascontiguousarray(clip_off, dtype=np.uint64)
        nq = clip_off.size - 1
        hits = np.zeros(nq, HIT_DTYPE)
        self._chk(lib().tir_group_search(self._g, _p(pcm), _p(clip_off), nq, coefs, float(tolerance), int(freq_ignore_low),
                                         int(freq_ignore_high), _p(hits)))
        return hits


def sqlite_insert_fingerprints(sqlite3_handle, context, uuid_text, vq):
    """tir_sqlite_insert_fingerprints without a context (CPU-only callers: it touches no GPU)."""
    vq = np.ascontiguousarray(vq, dtype=np.int32).reshape(-1, 2)
    rc = lib().tir_sqlite_insert_fingerprints(None, C.c_void_p(sqlite3_handle), context.encode(), uuid_text.encode(), _p(vq),
                                              vq.shape[0])
    if rc != OK:
        raise TirError(rc, "tir_sqlite_insert_fingerprints failed")


class Stream:
    """tir_stream wrapper: feed slinear chunks as they arrive, finish() searches the recording."""

    def __init__(self, ctx):
        self._ctx = ctx
        self._s = C.c_void_p()
        ctx._chk(lib().tir_stream_open(ctx._h, C.byref(self._s)))

    def feed(self, chunk):
        chunk = np.ascontiguousarray(chunk, dtype=np.int16)
        self._ctx._chk(lib().tir_stream_feed(self._s, _p(chunk), chunk.size))

    @property
    def samples(self):
        return int(lib().tir_stream_samples(self._s))

    @property
    def frames_done(self):
        """frames already extracted on the device (while the recording is still being fed)"""
        return int(lib().tir_stream_frames_done(self._s))

    def finish(self, coefs=1, tolerance=0.001, freq_ignore_low=-1, freq_ignore_high=-1):
        hit = np.zeros(1, HIT_DTYPE)
        self._ctx._chk(lib().tir_stream_finish(self._s, coefs, float(tolerance), int(freq_ignore_low), int(freq_ignore_high),
                                               _p(hit)))
        return hit[0]

    def close(self):
        if self._s:
            lib().tir_stream_close(self._s)
            self._s = C.c_void_p()

    __del__ = close
