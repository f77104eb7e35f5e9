"""Timing of the shared-window match chain only (no parity): 10 M fingerprints x 94 frames, Q queries.
python tools/gpu_time_match.py [n_fingerprints] [Q]   -- prints ms per batch and queries/s; with
TIR_NCU=1 runs a few batches only (for an ncu launch list: ncu -k regex:tir_ --metrics gpu__time_duration.sum)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from asterisk_tiresias_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
COEFS = int(sys.argv[3]) if len(sys.argv) > 3 else 1
TOL = float(sys.argv[4]) if len(sys.argv) > 4 else 0.001
F = 94
dev = "cuda"
st = torch.cuda.Stream()
ctx = capi.Context(device=0, stream=st.cuda_stream)
g = torch.Generator(device=dev); g.manual_seed(991)
rows = n * F
uu = torch.randint(0, 256, (n, 16), dtype=torch.uint8, device=dev, generator=g)
v1 = torch.randint(15_500_000, 18_500_000, (rows,), dtype=torch.int32, device=dev, generator=g)
v2 = torch.randint(-5_000_000, 20_000_000, (rows,), dtype=torch.int32, device=dev, generator=g)
row_off = torch.arange(n + 1, device=dev, dtype=torch.int64) * F
torch.cuda.synchronize()
t0 = time.time()
ctx.db_load_dev(n, uu.data_ptr(), row_off.data_ptr(), v1.data_ptr(), v2.data_ptr(), rows)
print(f"index build {time.time() - t0:.2f} s")
gq = torch.Generator(device=dev); gq.manual_seed(4242)
qv = torch.randint(15_500_000, 18_500_000, (Q, F), dtype=torch.int32, device=dev, generator=gq).double() * 1e-6
qv[: Q // 10] = v1.view(n, F)[: Q // 10].double() * 1e-6
q2 = torch.rand((Q, F), device=dev, generator=gq, dtype=torch.float64) * 25.0 - 5.0
coef = torch.stack([torch.pow(10.0, qv / 10.0).float(), torch.pow(10.0, q2 / 10.0).float()], dim=2).contiguous()
foff = np.arange(Q + 1, dtype=np.uint64) * F
d_hits = torch.zeros(Q * 24, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
steps = 3 if os.environ.get("TIR_NCU") else (200 if COEFS == 1 else 10)
with torch.cuda.stream(st):
    for _ in range(3):
        ctx.match_dev(coef.data_ptr(), foff, d_hits.data_ptr(), COEFS, TOL)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        ctx.match_dev(coef.data_ptr(), foff, d_hits.data_ptr(), COEFS, TOL)
    e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
hits = d_hits.cpu().numpy().view(capi.HIT_DTYPE)
print(f"n={n} Q={Q} coefs={COEFS} tol={TOL}: {ms * 1e3:.1f} us per batch, {Q / ms * 1e3 / 1e6:.2f} M queries/s, found {(hits['match_count'] > 0).sum()}, "
      f"exact copies found with 94 votes: {(hits['match_count'][: Q // 10] == 94).sum()}/{Q // 10}")
