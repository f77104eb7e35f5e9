"""BASELINE config[3] extraction leg: 16 kHz audio with win 1024 / hop 512 (and 16 kHz with 512/256): kernel timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from asterisk_tiresias_b200 import capi
st = torch.cuda.Stream()
for win, sr in ((1024, 16000), (512, 16000), (512, 8000)):
    hop = win // 2
    ctx = capi.Context(device=0, win=win, hop=hop, samplerate=sr, stream=st.cuda_stream)
    n_clips, n = 2000, 30 * sr
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
        for _ in range(5):
            ctx.extract_dev(d_pcm.data_ptr(), off, d_coef.data_ptr(), d_vq.data_ptr())
        e1.record(st)
    st.synchronize()
    ms = e0.elapsed_time(e1) / 5
    bpf = hop * 2 + 16
    print(f"win {win} sr {sr}: {ms:.3f} ms, {F / ms * 1e3 / 1e9:.3f} G frames/s, {n_clips * 30 / ms * 1e3 / 1e6:.1f} M audio-s/s, {F * bpf / ms / 1e6:.0f} GB/s algorithmic")
    ctx.close(); del d_pcm, d_coef, d_vq
