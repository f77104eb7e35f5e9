"""Kernel timing only (no parity): python tools/gpu_time_extract.py [lib.so ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import subprocess
libs = sys.argv[1:] or [None]
if len(libs) > 1 or libs[0]:
    for l in libs:  # one process per library (the binding caches the handle)
        if l:
            r = subprocess.run([sys.executable, __file__], env=dict(os.environ, TIR_LIB=l), capture_output=True, text=True)
            print(os.path.basename(l), r.stdout.strip() or r.stderr[-300:])
    sys.exit(0)
from asterisk_tiresias_b200 import capi
if os.environ.get("TIR_LIB"):
    capi.LIB_PATH = os.environ["TIR_LIB"]
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
print(f"ms={ms:.4f} frames/s={F / ms * 1e3 / 1e9:.4f}G cycles/frame/SM={ms * 1e-3 * 1.93e9 / (F / 148):.1f}")
