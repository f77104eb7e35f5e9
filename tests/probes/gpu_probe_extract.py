"""Quick GPU probe: extraction parity vs the oracle on a small corpus + a rough kernel timing."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from asterisk_tiresias_b200 import capi, synth
from oracle import pyoracle as po

out = {}
torch.cuda.init()
st = torch.cuda.Stream()
ctx = capi.Context(device=0, stream=st.cuda_stream)
plan = po.Plan()
for name, kw in [("uniform3s", dict(n_clips=24, seconds=3.0)), ("ragged", dict(n_clips=40, seconds=2.0, ragged=True)),
                 ("ulaw", dict(n_clips=8, seconds=3.0, ulaw=True))]:
    pcm, off = synth.make_corpus(**kw)
    if name == "ragged":  # odd sample counts -> unaligned clip starts
        lens = np.diff(off.astype(np.int64)); lens = lens - (np.arange(lens.size) % 7)
        clips = [pcm[int(off[i]):int(off[i]) + int(lens[i])] for i in range(lens.size)]
        pcm = np.concatenate(clips); off = np.zeros(lens.size + 1, np.uint64); off[1:] = np.cumsum(lens)
    coef, vq = ctx.extract(pcm, off)
    oc, oy, ov = plan.extract_batch(pcm, off, n_threads=8)
    same = (coef.view(np.uint32) == oc.view(np.uint32))
    out[name] = dict(frames=int(coef.shape[0]), coef_bit_identical=float(same.mean()), vq_identical=float((vq == ov).mean()))
    print(name, out[name], flush=True)

# timing: device-resident random PCM, 2000 clips x 30 s
n_clips, n = 2000, 240000
g = torch.Generator(device="cuda"); g.manual_seed(1)
d_pcm = torch.randint(-20000, 20000, (n_clips * n,), dtype=torch.int16, device="cuda", generator=g)
off = (np.arange(n_clips + 1, dtype=np.uint64) * n)
F = ctx.n_frames(off)
d_coef = torch.empty((F, 2), dtype=torch.float32, device="cuda")
d_vq = torch.empty((F, 2), dtype=torch.int32, device="cuda")
with torch.cuda.stream(st):
    for _ in range(3):
        ctx.extract_dev(d_pcm.data_ptr(), off, d_coef.data_ptr(), d_vq.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    reps = 5
    for _ in range(reps):
        ctx.extract_dev(d_pcm.data_ptr(), off, d_coef.data_ptr(), d_vq.data_ptr())
    e1.record(st)
st.synchronize()
ms = e0.elapsed_time(e1) / reps
out["timing"] = dict(ms=ms, frames=F, frames_per_s=F / ms * 1e3, audio_s_per_s=n_clips * 30 / ms * 1e3,
                     gbs=F * 528 / ms / 1e6)
print(out["timing"])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_extract.json", "w"), indent=1)
