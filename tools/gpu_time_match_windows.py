"""How the match chain's time depends on the number of DISTINCT windows of a batch (coefs = 1): the stored max1 values
are spread over `span` dB and every query frame takes one of K integer values, K = 4 .. 96.
python tools/gpu_time_match_windows.py [n_fingerprints] [Q]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from asterisk_tiresias_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
F = 94
SPAN = 96
dev = "cuda"
st = torch.cuda.Stream()
ctx = capi.Context(device=0, stream=st.cuda_stream)
g = torch.Generator(device=dev); g.manual_seed(991)
rows = n * F
uu = torch.randint(0, 256, (n, 16), dtype=torch.uint8, device=dev, generator=g)
v1 = torch.randint(-60_000_000, -60_000_000 + SPAN * 1_000_000, (rows,), dtype=torch.int32, device=dev, generator=g)
v2 = torch.randint(-5_000_000, 20_000_000, (rows,), dtype=torch.int32, device=dev, generator=g)
row_off = torch.arange(n + 1, device=dev, dtype=torch.int64) * F
ctx.db_load_dev(n, uu.data_ptr(), row_off.data_ptr(), v1.data_ptr(), v2.data_ptr(), rows)
foff = np.arange(Q + 1, dtype=np.uint64) * F
d_hits = torch.zeros(Q * 24, dtype=torch.uint8, device=dev)
gq = torch.Generator(device=dev); gq.manual_seed(4242)
for K, tol in ((4, 0.001), (4, 0.01), (4, 0.05), (11, 0.001), (16, 0.001), (32, 0.001), (33, 0.001), (48, 0.001), (64, 0.001), (64, 0.01), (65, 0.001), (96, 0.001)):
    qi = torch.randint(0, K, (Q, F), device=dev, generator=gq)
    qv = (qi - 59).double() + 0.5 * torch.sign((qi - 59).double())      # trunc() gives the integer back
    q2 = torch.zeros((Q, F), device=dev, dtype=torch.float64)
    coef = torch.stack([torch.pow(10.0, qv / 10.0).float(), torch.pow(10.0, q2 / 10.0).float()], dim=2).contiguous()
    torch.cuda.synchronize()
    steps = 50 if K <= 64 else 5
    with torch.cuda.stream(st):
        for _ in range(3):
            ctx.match_dev(coef.data_ptr(), foff, d_hits.data_ptr(), 1, tol)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(steps):
            ctx.match_dev(coef.data_ptr(), foff, d_hits.data_ptr(), 1, tol)
        e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    hits = d_hits.cpu().numpy().view(capi.HIT_DTYPE)
    print(f"K={K:3d} tol={tol}: {ms * 1e3:9.1f} us per batch of {Q}, {Q / ms * 1e3 / 1e6:8.3f} M queries/s, "
          f"best match_count {hits['match_count'].max()}", flush=True)
