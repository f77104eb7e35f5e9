import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from asterisk_tiresias_b200 import capi, synth
os.environ["TIR_DEBUG"] = "1"
devs = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "0,0,0").split(",")]
pcm, off = synth.make_corpus(100, 2.0, first_index=20000, ragged=True)
ref = capi.Context(device=0)
grp = capi.Group(devs)
coef, vq = ref.extract(pcm, off)
fo = np.concatenate([[0], np.cumsum((np.diff(off.astype(np.int64)) + 255) // 256)]).astype(np.uint64)
uu = np.stack([capi.uuid_to_bytes(synth.uuid_for(930000 + i)) for i in range(100)])
ref.db_load(uu, fo, vq[:, 0], vq[:, 1]); grp.db_load(uu, fo, vq[:, 0], vq[:, 1])
q_clips = [pcm[int(off[i]):int(off[i + 1])] for i in range(0, 100, 7)] + [np.zeros(0, np.int16)]
qoff = np.zeros(len(q_clips) + 1, np.uint64); qoff[1:] = np.cumsum([c.size for c in q_clips])
qpcm = np.concatenate(q_clips)
for it in range(4):
    for coefs, tol in ((1, 0.001), (1, 0.05), (2, 0.8)):
        t0 = time.time()
        try:
            g = grp.search(qpcm, qoff, coefs, tol)
            r = ref.search(qpcm, qoff, coefs, tol)
            print(it, coefs, tol, "ok", bool((g["match_count"] == r["match_count"]).all()), f"{time.time() - t0:.3f}s", grp.stats(), flush=True)
        except Exception as e:
            print(it, coefs, tol, "FAIL", e, f"{time.time() - t0:.3f}s", flush=True)
