"""EXPERIMENT: the extraction kernel built with -DTIR_RELAXED (tools/relaxed_experiment.sh: MUFU square
root and logarithm, contracted multiply-adds) against the product kernel -- speed on 2 000 x 30 s clips, and the gates of
the north star on 500 clips of the bench corpus: MFCC relative error, identical micro-unit hashes, identical trunc(max1)
(the window a query frame asks for), identical window membership of stored frames at tolerance 0.001.
One process per library (the binding caches the handle).  python tools/gpu_relaxed_experiment.py"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

RELAXED = os.path.join(ROOT, "tools", "_build", "libtiresias_gpu_relaxed.so")
OUT = os.path.join(ROOT, "gpurun_out")


def child(tag):
    import torch
    from asterisk_tiresias_b200 import capi, synth
    if tag == "relaxed":
        capi.LIB_PATH = RELAXED
        import asterisk_tiresias_b200.build as b
        b.needs_build = lambda: False
    st = torch.cuda.Stream()
    ctx = capi.Context(device=0, stream=st.cuda_stream)
    n_clips, n = 2000, 240000
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    d_pcm = torch.randint(-20000, 20000, (n_clips * n,), dtype=torch.int16, device="cuda", generator=g)
    off = (np.arange(n_clips + 1, dtype=np.uint64) * n)
    F = ctx.n_frames(off)
    d_coef = torch.empty((F, 2), dtype=torch.float32, device="cuda"); d_vq = torch.empty((F, 2), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        for _ in range(3):
            ctx.extract_dev(d_pcm.data_ptr(), off, d_coef.data_ptr(), d_vq.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(10):
            ctx.extract_dev(d_pcm.data_ptr(), off, d_coef.data_ptr(), d_vq.data_ptr())
        e1.record(st)
    st.synchronize()
    ms = e0.elapsed_time(e1) / 10
    pcm, coff = synth.make_corpus(500, 30.0, first_index=0)          # tone / noise / chirp clips, as in the bench
    coef, vq = ctx.extract(pcm, coff)
    np.savez(os.path.join(OUT, f"relaxed_exp_{tag}.npz"), coef=coef, vq=vq)
    print(json.dumps({"lib": tag, "ms_per_2000_clips": ms, "frames_per_s": F / ms * 1e3, "audio_s_per_s": n_clips * 30.0 / ms * 1e3}))


if len(sys.argv) > 1:
    child(sys.argv[1])
    sys.exit(0)
os.makedirs(OUT, exist_ok=True)
res = {}
for tag in ("exact", "relaxed"):
    r = subprocess.run([sys.executable, __file__, tag], capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not line:
        print(tag, "failed:", r.stderr[-600:]); sys.exit(1)
    res[tag] = json.loads(line[-1])
a, b = np.load(os.path.join(OUT, "relaxed_exp_exact.npz")), np.load(os.path.join(OUT, "relaxed_exp_relaxed.npz"))
ca, cb, va, vb = a["coef"].astype(np.float64), b["coef"].astype(np.float64), a["vq"].astype(np.int64), b["vq"].astype(np.int64)
rel = np.abs(cb - ca) / np.maximum(np.abs(ca), 1e-30)
ya, yb = va[:, 0] * 1e-6, vb[:, 0] * 1e-6
ta, tb = np.trunc(ya), np.trunc(yb)
near = lambda v: np.abs(v - np.round(v / 1e6) * 1e6) <= 1000          # a stored max1 inside SOME query window at tol 0.001
out = {"speed": res, "speedup": res["relaxed"]["frames_per_s"] / res["exact"]["frames_per_s"], "frames": int(ca.shape[0]),
       "mfcc_rel_err_max": float(rel.max()), "mfcc_rel_err_p999": float(np.quantile(rel, 0.999)), "mfcc_rel_err_mean": float(rel.mean()),
       "mfcc_within_1e-4": float((rel <= 1e-4).mean()),
       "hash_identical": float((va == vb).all(axis=1).mean()), "trunc_max1_identical": float((ta == tb).mean()),
       "window_membership_identical_tol_0.001": float((near(va[:, 0]) == near(vb[:, 0])).mean()),
       "note": "relaxed = -DTIR_RELAXED build of the same kernel (sqrt.approx, lg2.approx, contracted multiply-adds) against the product (bit-identical to the oracle)"}
print(json.dumps(out))
json.dump(out, open(os.path.join(OUT, "relaxed_experiment.json"), "w"), indent=1)
for t in ("exact", "relaxed"):
    os.remove(os.path.join(OUT, f"relaxed_exp_{t}.npz"))
