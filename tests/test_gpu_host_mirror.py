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


def test_directory_sync_batches_new_files_and_drops_removed_ones(oracle, tmp_path):
    """init_audio() (src/app_tiresias.c:324-551): what module load does with the configured directories."""
    plan = oracle.Plan()
    d1, d2 = tmp_path / "prompts", tmp_path / "wide"
    d1.mkdir(); d2.mkdir()
    clips = {}
    for i in range(40):
        pcm = synth.make_clip(9100 + i, 1.0 + (i % 5) * 0.5)
        fp.write_wav(str(d1 / f"p{i:02d}.wav"), pcm)
        clips[f"p{i:02d}.wav"] = (pcm, 8000)
    for i in range(6):                                          # a second context at 16 kHz: another plan, same table
        pcm = synth.make_clip(9200 + i, 1.5, samplerate=16000)
        fp.write_wav(str(d2 / f"w{i}.wav"), pcm, rate=16000)
        clips[f"w{i}.wav"] = (pcm, 16000)
    fp.write_wav(str(d1 / "copy-of-p00.wav"), clips["p00.wav"][0])   # same bytes, other name: one audio (P7);
    clips["copy-of-p00.wav"] = clips["p00.wav"]                        # alphasort lists the copy first
    assert fp.fp_init(None, 0)
    try:
        assert fp.fp_create_context_list_info("prompts", str(d1), False) and fp.fp_create_context_list_info("wide", str(d2), False)
        assert fp.fp_sync_directories() == 46
        assert fp.fp_sync_directories() == 0                      # idempotent
        lst = fp.fp_get_audio_lists_all()
        assert len(lst) == 46 and len(fp.fp_get_audio_lists_by_contextname("wide")) == 6
        sq, by_uuid = oracle.SqliteDB(), {}
        plans = {8000: plan, 16000: oracle.Plan(samplerate=16000)}
        for a in lst:
            pcm, sr = clips[a["name"]]
            sq.add_audio(a["uuid"], plans[sr].extract(pcm)[1], context=a["context"], name=a["name"])
            by_uuid[a["uuid"]] = a["name"]
        for name in ("p07.wav", "p39.wav", "w3.wav"):
            pcm, sr = clips[name]
            for tol in (0.01, 0.2):
                r = fp.fp_search_fingerprint_info("prompts", str((d2 if sr == 16000 else d1) / name), 1, tol)
                assert got(r) == expect(sq, plans[sr], pcm, by_uuid, 1, tol), (name, tol)
        # files removed from / added to the directory between two loads
        os.remove(d1 / "p07.wav"); os.remove(d1 / "p08.wav")
        extra = synth.make_clip(9300, 2.0)
        fp.write_wav(str(d1 / "new.wav"), extra)
        assert fp.fp_sync_directories() == 1
        names = sorted(a["name"] for a in fp.fp_get_audio_lists_by_contextname("prompts"))
        assert "p07.wav" not in names and "p08.wav" not in names and "new.wav" in names and len(names) == 39
        now = {a["name"]: a["uuid"] for a in fp.fp_get_audio_lists_by_contextname("prompts")}
        for a in lst:
            if a["name"] in ("p07.wav", "p08.wav"):
                sq.delete_audio(a["uuid"])
        sq.add_audio(now["new.wav"], plan.extract(extra)[1], context="prompts", name="new.wav")
        by_uuid[now["new.wav"]] = "new.wav"
        for pcm in (extra, clips["p07.wav"][0], clips["p20.wav"][0]):
            fp.write_wav(str(tmp_path / "q.wav"), pcm)
            for tol in (0.001, 0.3):
                assert got(fp.fp_search_fingerprint_info("prompts", str(tmp_path / "q.wav"), 1, tol)) == expect(sq, plan, pcm, by_uuid, 1, tol)
    finally:
        fp.fp_term()


def test_stereo_files_are_fingerprinted_from_the_channel_mean(oracle, tmp_path):
    """A two-channel WAV (src/fp_handler.c:604: aubio opens whatever the file is and averages the channels in float):
    fingerprinted, stored and searched like the oracle chain run on the float mean."""
    plan = oracle.Plan()
    assert fp.fp_init(None, 0)
    try:
        rng = np.random.default_rng(5)
        sq, by_uuid, files = oracle.SqliteDB(), {}, {}
        for i in range(6):
            l = synth.make_clip(300 + i, 3.0).astype(np.int32)
            r = np.clip(l // 2 + rng.integers(-2000, 2000, l.size), -32768, 32767)
            st = np.stack([l, r], axis=1).astype(np.int16)
            name = f"st-{i}.wav"
            fp.write_wav(str(tmp_path / name), st.reshape(-1), channels=2)
            files[name] = st
            assert fp.fp_craete_audio_list_info("c", str(tmp_path / name))
        for a in fp.fp_get_audio_lists_by_contextname("c"):
            _, y, _ = plan.extract_interleaved(files[a["name"]], 2)
            sq.add_audio(a["uuid"], y, context="c", name=a["name"])
            by_uuid[a["uuid"]] = a["name"]
        for name in ("st-0.wav", "st-4.wav"):
            for coefs, tol in ((1, 0.001), (2, 0.5)):
                _, y, _ = plan.extract_interleaved(files[name], 2)
                h = sq.search(y, coefs, tol, -1, -1, has_y=np.isfinite(y))
                want = None if h is None else (h["uuid"], by_uuid[h["uuid"]], h["match_count"], h["frame_count"])
                assert got(fp.fp_search_fingerprint_info("c", str(tmp_path / name), coefs, tol)) == want
    finally:
        fp.fp_term()


def test_non_pcm16_files_are_refused(tmp_path):
    assert fp.fp_init(None, 0)
    try:
        open(tmp_path / "junk.wav", "wb").write(b"not a wav file at all")
        assert not fp.fp_craete_audio_list_info("c", str(tmp_path / "junk.wav"))
        assert fp.fp_search_fingerprint_info("c", str(tmp_path / "junk.wav")) is None
        u = fp.fp_generate_uuid()
        assert len(u) == 36 and u[14] == "4" and u.count("-") == 4
    finally:
        fp.fp_term()
