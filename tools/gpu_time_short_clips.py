"""Extraction of many short clips (SURVEY 8d config 2, 3 s variant): step time vs kernel time.
python tools/gpu_time_short_clips.py [n_clips] [seconds]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from asterisk_tiresias_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
sec = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
ns = int(8000 * sec)
st = torch.cuda.Stream()
ctx = capi.Context(device=0, stream=st.cuda_stream)
ctx.set_profiling(True)
g = torch.Generator(device="cuda"); g.manual_seed(1)
d_pcm = torch.randint(-20000, 20000, (n * ns,), dtype=torch.int16, device="cuda", generator=g)
off = np.arange(n + 1, dtype=np.uint64) * ns
F = ctx.n_frames(off)
d_coef = torch.empty((F, 2), dtype=torch.float32, device="cuda"); d_vq = torch.empty((F, 2), dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
with torch.cuda.stream(st):
    for _ in range(6):
        ctx.extract_dev(d_pcm.data_ptr(), off, d_coef.data_ptr(), d_vq.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record(st)
    for _ in range(5):
        ctx.extract_dev(d_pcm.data_ptr(), off, d_coef.data_ptr(), d_vq.data_ptr())
    e1.record(st)
    t1 = time.time()
    torch.cuda.synchronize()
print(f"{n} clips x {sec} s: {e0.elapsed_time(e1) / 5:.3f} ms per step (host enqueue {(t1 - t0) / 5 * 1e3:.3f} ms), extraction kernel {ctx.last_kernel_ms(0):.3f} ms, "
      f"{n * sec / (e0.elapsed_time(e1) / 5 * 1e-3) / 1e6:.2f} M audio-s/s")
