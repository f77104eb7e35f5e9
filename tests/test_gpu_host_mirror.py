"""BASELINE config[0] through the host-side mirror of src/fp_handler.h (libtiresias_host.so): a
directory of WAV files is fingerprinted into the SQLite database + device table, recordings are
searched, audios deleted, the database backed up and restored -- every search result equal to the
oracle chain (oracle extraction -> reference SQL text on the real SQLite)."""
import hashlib
import os

import numpy as np
import pytest

from asterisk_tiresias_b200 import synth
from asterisk_tiresias_b200.host import fp_host as fp

pytestmark = pytest.mark.gpu


def expect(sq, plan, pcm, by_uuid, coefs=1, tol=0.001, lo=-1, hi=-1):
    _, y, _ = plan.extract(pcm)
    h = sq.search(y, coefs, tol, lo, hi, has_y=np.isfinite(y))
    if h is None:
        return None
    return (h["uuid"], by_uuid[h["uuid"]], h["match_count"], h["frame_count"])


def got(r):
    return None if r is None else (r["uuid"], r["name"], r["match_count"], r["frame_count"])


def test_directory_fingerprint_search_delete_backup_restore(oracle, tmp_path):
    plan = oracle.Plan()
    backup = str(tmp_path / "audio_recongition.db")          # [sic] DEF_BACKUP_DATABASE, src/fp_handler.c:31
    assert fp.fp_init(backup, 0)
    try:
        assert fp.fp_create_context_list_info("ivr", str(tmp_path), False)
        clips = {}
        for i in range(100):                                   # a 100-entry DB; entry 0 is the 30 s clip of config[0]
            pcm = synth.make_clip(7000 + i, 30.0 if i == 0 else 3.0 + (i % 4), ulaw=(i % 2 == 0))
            name = f"prompt-{i:03d}.wav"
            fp.write_wav(str(tmp_path / name), pcm)
            clips[name] = pcm
            assert fp.fp_craete_audio_list_info("ivr", str(tmp_path / name))
        # P7: the same file again (same context, same md5) is skipped and reported as success
        assert fp.fp_craete_audio_list_info("ivr", str(tmp_path / "prompt-003.wav"))
        lst = fp.fp_get_audio_lists_by_contextname("ivr")
        assert len(lst) == 100 and fp.fp_get_audio_lists_by_contextname("nope") == []
        a3 = next(a for a in lst if a["name"] == "prompt-003.wav")
        assert a3["hash"] == hashlib.md5(open(tmp_path / "prompt-003.wav", "rb").read()).hexdigest() == fp.fp_create_hash(str(tmp_path / "prompt-003.wav"))
        assert not fp.fp_craete_audio_list_info("ivr", str(tmp_path / "missing.wav"))
        # the oracle chain with the uuids the module generated
        sq = oracle.SqliteDB()
        by_uuid = {}
        for a in lst:
            _, y, _ = plan.extract(clips[a["name"]])
            sq.add_audio(a["uuid"], y, context="ivr", name=a["name"])
            by_uuid[a["uuid"]] = a["name"]
        queries = [("prompt-000.wav", 1, 0.001), ("prompt-017.wav", 1, 0.001), ("prompt-042.wav", 1, 0.05), ("prompt-099.wav", 2, 0.5)]
        for name, coefs, tol in queries:
            r = fp.fp_search_fingerprint_info("ivr", str(tmp_path / name), coefs, tol)
            assert got(r) == expect(sq, plan, clips[name], by_uuid, coefs, tol), name
        # the context argument is not part of the match (P2)
        assert got(fp.fp_search_fingerprint_info("other", str(tmp_path / "prompt-017.wav"))) == expect(sq, plan, clips["prompt-017.wav"], by_uuid)
        # an unknown recording, freq_ignore on, and the argument rules
        stranger = synth.make_clip(424242, 3.0)
        fp.write_wav(str(tmp_path / "q.wav"), stranger)
        for tol, lo, hi in ((0.001, -1, -1), (0.3, -1, -1), (0.3, 40, 70)):
            assert got(fp.fp_search_fingerprint_info("ivr", str(tmp_path / "q.wav"), 1, tol, lo, hi)) == expect(sq, plan, stranger, by_uuid, 1, tol, lo, hi)
        assert fp.fp_search_fingerprint_info("ivr", str(tmp_path / "q.wav"), 3, 0.001) is None       # src/fp_handler.c:247
        assert fp.fp_search_fingerprint_info("ivr", str(tmp_path / "missing.wav")) is None
        # delete two audios (CLI "tiresias remove audio <uuid>"), both stores follow
        for name in ("prompt-017.wav", "prompt-000.wav"):
            u = next(a["uuid"] for a in lst if a["name"] == name)
            assert fp.fp_delete_audio_list_info(u)
            assert not fp.fp_delete_audio_list_info(u)                                                 # already gone
            sq.delete_audio(u)
        for name, coefs, tol in queries:
            assert got(fp.fp_search_fingerprint_info("ivr", str(tmp_path / name), coefs, tol)) == expect(sq, plan, clips[name], by_uuid, coefs, tol)
        # unload / load: the backup file restores SQLite, which restores the device table
        assert fp.fp_term() and os.path.getsize(backup) > 0
        assert fp.fp_init(backup, 0)
        assert len(fp.fp_get_audio_lists_all()) == 98
        for name, coefs, tol in queries:
            assert got(fp.fp_search_fingerprint_info("ivr", str(tmp_path / name), coefs, tol)) == expect(sq, plan, clips[name], by_uuid, coefs, tol)
        # "tiresias remove context <ctx>": its audios, their fingerprints, then the context
        assert fp.fp_delete_context_list_info("ivr") and not fp.fp_delete_context_list_info("ivr")
        assert fp.fp_get_audio_lists_all() == []
        assert fp.fp_search_fingerprint_info("ivr", str(tmp_path / "prompt-042.wav")) is None
    finally:
        fp.fp_term()


def test_non_mono_or_non_pcm16_files_are_refused(tmp_path):
    assert fp.fp_init(None, 0)
    try:
        pcm = synth.make_clip(1, 1.0)
        fp.write_wav(str(tmp_path / "stereo.wav"), np.repeat(pcm, 2), channels=2)
        open(tmp_path / "junk.wav", "wb").write(b"not a wav file at all")
        assert not fp.fp_craete_audio_list_info("c", str(tmp_path / "stereo.wav"))
        assert fp.fp_search_fingerprint_info("c", str(tmp_path / "junk.wav")) is None
        u = fp.fp_generate_uuid()
        assert len(u) == 36 and u[14] == "4" and u.count("-") == 4
    finally:
        fp.fp_term()
