"""coefs=2 (per-query path) timing + brute-force check on a mid-size DB."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from asterisk_tiresias_b200 import capi
dev = torch.device("cuda", 0)
ctx = capi.Context(device=0, stream=torch.cuda.current_stream().cuda_stream)
ctx.set_profiling(True)
n, F, Q = 2_000_000, 94, 100
g = torch.Generator(device=dev); g.manual_seed(5)
uu = torch.randint(0, 256, (n, 16), dtype=torch.uint8, device=dev, generator=g)
v1 = torch.randint(15_500_000, 18_500_000, (n * F,), dtype=torch.int32, device=dev, generator=g)
v2 = torch.randint(-5_000_000, 20_000_000, (n * F,), dtype=torch.int32, device=dev, generator=g)
off = torch.arange(n + 1, device=dev, dtype=torch.int64) * F
ctx.db_load_dev(n, uu.data_ptr(), off.data_ptr(), v1.data_ptr(), v2.data_ptr(), n * F)
y1 = torch.randint(15_500_000, 18_500_000, (Q, F), device=dev, generator=g).double() * 1e-6
y2 = torch.rand((Q, F), device=dev, generator=g, dtype=torch.float64) * 25 - 5
coef = torch.stack([torch.pow(10.0, y1 / 10).float(), torch.pow(10.0, y2 / 10).float()], dim=2).contiguous()
foff = np.arange(Q + 1, dtype=np.uint64) * F
d_hits = torch.zeros(Q * 24, dtype=torch.uint8, device=dev)
for coefs, tol in ((2, 0.001), (2, 0.5), (1, 0.001)):
    for _ in range(2):
        ctx.match_dev(coef.data_ptr(), foff, d_hits.data_ptr(), coefs, tol)
    torch.cuda.synchronize()
    t = time.time()
    for _ in range(3):
        ctx.match_dev(coef.data_ptr(), foff, d_hits.data_ptr(), coefs, tol)
    torch.cuda.synchronize()
    ms = (time.time() - t) / 3 * 1e3
    hits = d_hits.cpu().numpy().view(capi.HIT_DTYPE)
    # brute force for 3 queries
    yy1 = 10 * torch.log10(coef[:, :, 0].double()); yy2 = 10 * torch.log10(coef[:, :, 1].double())
    ok = 0
    for q in (0, 1, Q - 1):
        votes = torch.zeros(n, dtype=torch.int32, device=dev)
        for f in range(F):
            k = int(torch.trunc(yy1[q, f]).item())
            lo1, hi1 = round((k - tol) * 1e6), round((k + tol) * 1e6)
            m = (v1 >= lo1) & (v1 <= hi1)
            if coefs == 2:
                c = float(yy2[q, f].item())
                m &= (v2 >= round((c - tol) * 1e6)) & (v2 <= round((c + tol) * 1e6))
            votes += m.view(n, F).any(dim=1).to(torch.int32)
        best = int(votes.max().item())
        ok += int(hits["match_count"][q] == best)
    print(f"coefs={coefs} tol={tol}: {ms:.3f} ms per {Q} queries (kernel {ctx.last_kernel_ms(1):.3f} ms), found {int((hits['match_count']>0).sum())}, max count {int(hits['match_count'].max())}, brute-force count agreement {ok}/3")
