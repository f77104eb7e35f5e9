"""ctypes stub over libtiresias_host.so (the host-side mirror of src/fp_handler.h).  Test / bench glue only."""
from __future__ import annotations

import ctypes as C
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "..", "libtiresias_host.so")


class AudioInfo(C.Structure):
    _fields_ = [("uuid", C.c_char * 40), ("name", C.c_char * 256), ("context", C.c_char * 256), ("hash", C.c_char * 40),
                ("frame_count", C.c_int), ("match_count", C.c_int)]

    def as_dict(self):
        return {"uuid": self.uuid.decode(), "name": self.name.decode(), "context": self.context.decode(), "hash": self.hash.decode(),
                "frame_count": self.frame_count, "match_count": self.match_count}


class ContextInfo(C.Structure):
    _fields_ = [("name", C.c_char * 256), ("directory", C.c_char * 1024)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        from .. import build
        if build.needs_build() or not os.path.exists(LIB_PATH):
            build.build()
        L = C.CDLL(LIB_PATH)
        L.fp_init.restype = C.c_bool
        L.fp_init.argtypes = [C.c_char_p, C.c_int]
        L.fp_term.restype = C.c_bool
        L.fp_create_context_list_info.restype = C.c_bool
        L.fp_create_context_list_info.argtypes = [C.c_char_p, C.c_char_p, C.c_bool]
        L.fp_delete_context_list_info.restype = C.c_bool
        L.fp_delete_context_list_info.argtypes = [C.c_char_p]
        L.fp_get_context_lists_all.argtypes = [C.c_void_p, C.c_int]
        L.fp_get_audio_lists_all.argtypes = [C.c_void_p, C.c_int]
        L.fp_get_audio_lists_by_contextname.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
        L.fp_craete_audio_list_info.restype = C.c_bool
        L.fp_craete_audio_list_info.argtypes = [C.c_char_p, C.c_char_p]
        L.fp_delete_audio_list_info.restype = C.c_bool
        L.fp_delete_audio_list_info.argtypes = [C.c_char_p]
        L.fp_search_fingerprint_info.restype = C.c_bool
        L.fp_search_fingerprint_info.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_void_p]
        L.fp_create_hash.restype = C.c_void_p
        L.fp_create_hash.argtypes = [C.c_char_p]
        L.fp_generate_uuid.restype = C.c_void_p
        L.fp_sqlite_handle.restype = C.c_void_p
        _lib = L
    return _lib


def _enc(s):
    return None if s is None else s.encode()


def fp_init(backup_db=None, device=0):
    return bool(lib().fp_init(_enc(backup_db), device))


def fp_term():
    return bool(lib().fp_term())


def fp_create_context_list_info(name, directory, replace=False):
    return bool(lib().fp_create_context_list_info(_enc(name), _enc(directory), replace))


def fp_delete_context_list_info(name):
    return bool(lib().fp_delete_context_list_info(_enc(name)))


def fp_get_audio_lists_all(cap=100000):
    buf = (AudioInfo * cap)()
    n = lib().fp_get_audio_lists_all(buf, cap)
    return [buf[i].as_dict() for i in range(max(0, min(n, cap)))]


def fp_get_audio_lists_by_contextname(name, cap=100000):
    buf = (AudioInfo * cap)()
    n = lib().fp_get_audio_lists_by_contextname(_enc(name), buf, cap)
    return None if n < 0 else [buf[i].as_dict() for i in range(min(n, cap))]


def fp_craete_audio_list_info(context, filename):
    return bool(lib().fp_craete_audio_list_info(_enc(context), _enc(filename)))


def fp_delete_audio_list_info(uuid):
    return bool(lib().fp_delete_audio_list_info(_enc(uuid)))


def fp_search_fingerprint_info(context, filename, coefs=1, tolerance=0.001, freq_ignore_low=-1, freq_ignore_high=-1):
    """-> dict (uuid, name, context, hash, frame_count, match_count) or None (the reference's NULL)"""
    out = AudioInfo()
    ok = lib().fp_search_fingerprint_info(_enc(context), _enc(filename), coefs, tolerance, freq_ignore_low, freq_ignore_high, C.byref(out))
    return out.as_dict() if ok else None


def fp_sync_directories():
    return int(lib().fp_sync_directories())


def fp_create_hash(filename):
    p = lib().fp_create_hash(_enc(filename))
    if not p:
        return None
    s = C.string_at(p).decode()
    C.CDLL(None).free(C.c_void_p(p))
    return s


def fp_generate_uuid():
    p = lib().fp_generate_uuid()
    s = C.string_at(p).decode()
    C.CDLL(None).free(C.c_void_p(p))
    return s


def write_wav(path, pcm, rate=8000, channels=1):
    """44-byte canonical RIFF/WAVE PCM16 header (what ast_writefile(..., "wav") produces)."""
    pcm = np.ascontiguousarray(pcm, dtype="<i2")
    data = pcm.tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE" + b"fmt " +
                struct.pack("<IHHIIHH", 16, 1, channels, rate, rate * channels * 2, channels * 2, 16) +
                b"data" + struct.pack("<I", len(data)) + data)
