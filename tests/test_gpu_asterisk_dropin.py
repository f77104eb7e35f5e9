"""The Asterisk-facing surface, UNCHANGED, on the GPU path: the reference's module shell (src/app_tiresias.c),
dialplan application (src/application_handler.c) and CLI (src/cli_handler.c) -- compiled from /root/reference/src
as they are -- run on the replacement fp_handler.c inside a fake Asterisk (tests/fake_asterisk): module load scans
the configured directories and fingerprints them, Tiresias(context,duration,...) records from a scripted channel
and sets the TIR* channel variables, the four CLI commands list and delete.  Every result is compared with the
oracle chain: oracle extraction of the same files + the reference's SQL on the real SQLite."""
import os
import struct
import hashlib

import numpy as np
import pytest

from asterisk_tiresias_b200 import synth

pytestmark = pytest.mark.gpu
SR = 8000


def write_wav(path, pcm, rate=SR, channels=1):
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + pcm.nbytes) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, channels, rate, rate * 2 * channels, 2 * channels, 16))
        f.write(b"data" + struct.pack("<I", pcm.nbytes) + pcm.tobytes())


def parse_table(text, widths):
    rows = []
    for line in text.splitlines()[1:]:
        cols, at = [], 0
        for w in widths:
            cols.append(line[at:at + w].rstrip())
            at += w + 1
        rows.append(cols)
    return rows


STEREO_TWIN = 63001


@pytest.fixture(scope="module")
def world(tmp_path_factory, oracle):
    from fake_asterisk.harness import FakeAsterisk
    from fake_asterisk import build as fa_build
    if fa_build.build() is None:
        pytest.skip("drop-in module not built and /root/reference absent")
    root = str(tmp_path_factory.mktemp("fake_ast"))
    files = {}
    for ctx, first, n in (("music", 61000, 7), ("ivr", 62000, 5)):
        d = os.path.join(root, "audio", ctx)
        os.makedirs(d)
        for i in range(n):
            pcm = synth.make_clip(first + i, 4.0 + 0.37 * i, SR)
            name = f"{ctx}_{i:02d}.wav"
            write_wav(os.path.join(d, name), pcm)
            files[(ctx, name)] = pcm
    # two-channel files (aubio's source averages the channels in float): one with different channels, one whose channels
    # are both the clip STEREO_TWIN -- its mean is that clip, so a mono recording of it must find this file
    d = os.path.join(root, "audio", "st")
    os.makedirs(d)
    rng = np.random.default_rng(9)
    a = synth.make_clip(63000, 5.0, SR).astype(np.int32)
    twin = synth.make_clip(STEREO_TWIN, 4.0, SR)
    for name, st in (("st_00.wav", np.stack([a, np.clip(a // 3 + rng.integers(-1500, 1500, a.size), -32768, 32767)], axis=1).astype(np.int16)),
                     ("st_01.wav", np.repeat(twin[:, None], 2, axis=1))):
        write_wav(os.path.join(d, name), st.reshape(-1), channels=2)
        files[("st", name)] = st
    fa = FakeAsterisk(root)
    fa.write_conf(f"[global]\ntolerance=0.01\n\n[music]\ndirectory={root}/audio/music\n\n[ivr]\ndirectory = {root}/audio/ivr ; the announcements\n\n[st]\ndirectory={root}/audio/st\n")
    assert fa.load() == 0, fa.log()      # AST_MODULE_LOAD_SUCCESS: fp_init + directory scan + cli_init + application_init
    return {"fa": fa, "root": root, "files": files, "oracle": oracle}


def oracle_db(world, listing):
    """the oracle chain's database for the audios the module lists (uuid -> file)"""
    po = world["oracle"]
    plan = po.Plan(512, 256, 40, 2, SR)
    sq = po.SqliteDB()
    for uuid, name, ctx, _ in listing:
        x = world["files"][(ctx, name)]
        _, y, _ = plan.extract(x) if x.ndim == 1 else plan.extract_interleaved(x, x.shape[1])
        sq.add_audio(uuid, y, context=ctx, name=name)
    return plan, sq


def listing_of(fa, ctx):
    rc, out = fa.cli(f"tiresias show audios {ctx}")
    assert rc == 0
    assert out.splitlines()[0] == "%-36.36s %-45.45s %-36.36s %-36.36s" % ("Uuid", "Name", "Context", "Hash")   # src/cli_handler.c:132
    return [tuple(r) for r in parse_table(out, (36, 45, 36, 36))]


def test_module_load_fingerprints_the_directories_and_cli_lists_them(world):
    fa = world["fa"]
    assert fa.L.fake_cli_count() == 4
    rc, out = fa.cli("tiresias show contexts")
    assert rc == 0
    lines = out.splitlines()
    assert lines[0] == "%-36.36s %-70.70s" % ("Name", "Directory")                # src/cli_handler.c:78
    got = sorted(tuple(r) for r in parse_table(out, (36, 70)))
    assert got == sorted([("music", f"{world['root']}/audio/music"[:70]), ("ivr", f"{world['root']}/audio/ivr"[:70]), ("st", f"{world['root']}/audio/st"[:70])])
    for ctx, n in (("music", 7), ("ivr", 5), ("st", 2)):
        rows = listing_of(fa, ctx)
        assert len(rows) == n
        for uuid, name, c, h in rows:
            assert c == ctx and len(uuid) == 36
            assert h == hashlib.md5(open(os.path.join(world["root"], "audio", ctx, name), "rb").read()).hexdigest()   # src/fp_handler.c:758-805
    assert fa.cli("tiresias show audios")[0] == 1          # CLI_SHOWUSAGE: argc != 4
    assert fa.cli("tiresias show audios nosuch")[1].count("\n") == 1   # header only


def test_tiresias_application_sets_the_channel_variables_like_the_oracle_chain(world):
    fa = world["fa"]
    listing = listing_of(fa, "music") + listing_of(fa, "ivr") + listing_of(fa, "st")
    plan, sq = oracle_db(world, listing)
    by_uuid = {u: (n, c, h) for u, n, c, h in listing}
    cases = [
        ("st,3000,0.3", synth.make_clip(STEREO_TWIN, 4.0, SR), 3000, 0.3, -1, -1),                 # the mean of st_01.wav's channels
        ("music,3000", world["files"][("music", "music_03.wav")], 3000, 0.01, -1, -1),            # [global] tolerance applies
        ("ivr,2000,0.05", world["files"][("ivr", "ivr_01.wav")][4000:], 2000, 0.05, -1, -1),      # argument overrides it
        ("music,1500,0.5,1,5000", synth.make_clip(424242, 3.0, SR), 1500, 0.5, 1, 5000),          # unrelated audio, wide windows, freq_ignore_*
        ("music", world["files"][("music", "music_05.wav")], 3000, 0.01, -1, -1),                 # DEF_DURATION 3000 ms
    ]
    found = 0
    for data, audio, dur, tol, lo, hi in cases:
        rc, var, info = fa.exec_app(data, audio)
        assert rc == 0 and info["answered"] == 1                                                   # ast_answer on a channel that is not up
        n_rec = min(audio.size, dur * SR // 1000)
        assert info["samples_read"] == n_rec                                                       # 20 ms frames until `duration` ms have passed
        _, y, _ = plan.extract(audio[:n_rec])
        exp = sq.search(y, 1, tol, lo, hi, has_y=np.isfinite(y))                                   # coefs = 1: src/application_handler.c:180
        if exp is None:
            assert var == {"TIRSTATUS": "NOTFOUND"}, (data, var)
            continue
        found += 1
        name, ctx, h = by_uuid[exp["uuid"]]
        assert var == {"TIRSTATUS": "FOUND", "TIRFRAMECOUNT": str(exp["frame_count"]), "TIRMATCHCOUNT": str(exp["match_count"]),
                       "TIRFILEUUID": exp["uuid"], "TIRFILENAME": name, "TIRCONTEXT": ctx, "TIRFILEHASH": h}, (data, var, exp)
    assert found >= 3
    # a caller that hangs up after one second; a channel that is already up is not answered again
    rc, var, info = fa.exec_app("music,3000", world["files"][("music", "music_00.wav")], hangup_at=8000, state_up=True)
    assert rc == 0 and var == {"TIRSTATUS": "HANGUP"} and info["answered"] == 0
    # non-voice frames on the channel are skipped (src/application_handler.c:289-293); missing arguments are an error
    rc, var, _ = fa.exec_app("music,1000", world["files"][("music", "music_02.wav")], control_every=7)
    assert rc == 0 and var["TIRSTATUS"] in ("FOUND", "NOTFOUND")
    assert fa.exec_app("", np.zeros(10, np.int16))[0] == -1


def test_cli_remove_and_module_reload_follow_the_oracle_chain(world):
    fa = world["fa"]
    listing = listing_of(fa, "music") + listing_of(fa, "ivr") + listing_of(fa, "st")
    plan, sq = oracle_db(world, listing)
    audio = world["files"][("music", "music_03.wav")]
    _, y, _ = plan.extract(audio[:24000])
    first = sq.search(y, 1, 0.01, has_y=np.isfinite(y))
    assert first is not None
    # (the synthetic clips tie on match_count and the tie goes to the greatest uuid, which is random: the winner
    # may live in either context)
    removed_name, removed_ctx = {u: (n, c) for u, n, c, h in listing}[first["uuid"]]
    rc, out = fa.cli(f"tiresias remove audio {first['uuid']}")
    assert rc == 0 and out == f"Removed the audio info. uuid[{first['uuid']}]\n"                   # src/cli_handler.c:188
    sq.delete_audio(first["uuid"])
    rc, out = fa.cli(f"tiresias remove audio {first['uuid']}")
    assert rc == 2 and out == f"Could not remove the audio info. uuid[{first['uuid']}]\n"
    exp = sq.search(y, 1, 0.01, has_y=np.isfinite(y))
    rc, var, _ = fa.exec_app("music,3000", audio)
    assert (var["TIRSTATUS"] == "NOTFOUND") if exp is None else (var["TIRFILEUUID"] == exp["uuid"] and var["TIRMATCHCOUNT"] == str(exp["match_count"]))
    # remove a whole context: its audios go with it (src/fp_handler.c:1039-1095)
    rc, out = fa.cli("tiresias remove context ivr")
    assert rc == 0 and out == "Removed the context info. context[ivr]\n"
    for u, n, c, h in listing:
        if c == "ivr":
            sq.delete_audio(u)
    assert listing_of(fa, "ivr") == []
    q = world["files"][("ivr", "ivr_01.wav")]
    _, y2, _ = plan.extract(q[:24000])
    exp2 = sq.search(y2, 1, 0.05, has_y=np.isfinite(y2))
    rc, var2, _ = fa.exec_app("ivr,3000,0.05", q)
    assert (var2["TIRSTATUS"] == "NOTFOUND") if exp2 is None else (var2["TIRFILEUUID"] == exp2["uuid"] and var2["TIRMATCHCOUNT"] == str(exp2["match_count"]))
    # unload writes the backup database; load restores it (and the device table), re-scans the directories:
    # the file of the removed audio is still on disk, so it is fingerprinted again under a new uuid, `ivr` comes back
    before = sorted((n, c, h) for u, n, c, h in listing_of(fa, "music"))
    assert (removed_name, removed_ctx) not in [(n, c) for n, c, h in before]
    n_music = 7 - (removed_ctx == "music")
    assert len(before) == n_music
    assert fa.unload() == 0
    assert os.path.exists(os.environ["TIRESIAS_BACKUP_DATABASE"])
    assert fa.load() == 0, fa.log()
    after = listing_of(fa, "music")
    assert sorted((n, c, h) for u, n, c, h in after if (n, c) != (removed_name, removed_ctx)) == before
    assert len(after) == 7 and len(listing_of(fa, "ivr")) == 5, fa.log()[-3000:]
    assert any((n, c) == (removed_name, removed_ctx) for u, n, c, h in after + listing_of(fa, "ivr") + listing_of(fa, "st"))
    plan, sq = oracle_db(world, listing_of(fa, "music") + listing_of(fa, "ivr") + listing_of(fa, "st"))
    exp = sq.search(y, 1, 0.01, has_y=np.isfinite(y))
    rc, var, _ = fa.exec_app("music,3000", audio)
    assert exp is not None and var["TIRSTATUS"] == "FOUND" and var["TIRFILEUUID"] == exp["uuid"] and var["TIRMATCHCOUNT"] == str(exp["match_count"])
    assert fa.unload() == 0
