#!/usr/bin/env python
"""bench.py -- headline benchmark of the fingerprint hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Main line (one JSON line on stdout, rank 0):
  metric  audio_seconds_fingerprinted_per_second, BASELINE config[1]: batch MFCC extraction of
          10 000 synthetic 30 s 8 kHz clips (tone/noise/chirp/composite, G.711 mu-law round trip)
          per GPU.  A step = one pass of the fused extraction kernel over the whole batch.
  value   inputs resident in HBM when the timed region starts (tir_extract_dev)
  e2e     same metric through tir_extract with HOST (pinned) buffers: PCM host->device, kernel,
          coefficients + hashes device->host, every step
  roofline / cpu_baseline / clocks / gpu_launches : see DESIGN.md "Measurement"
  match   secondary object: match queries/s against the synthetic fingerprint DB, sharded by uuid
          over the N ranks (per-query top-1 crosses NVLink through an NCCL all-gather)

--impl reference: the reference's CPU path for the same metric -- the oracle restatement of
libaubio (oracle/, kind "port": the reference itself needs Asterisk + libaubio and cannot be built
here) on all host cores, each step a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR, SECONDS, WIN, HOP = 8000, 30.0, 512, 256
N_SAMP = int(SR * SECONDS)
FRAMES_PER_CLIP = -(-N_SAMP // HOP)            # 938
BYTES_PER_FRAME = HOP * 2 + 16                 # 528 B / frame algorithmic
BYTES_PER_AUDIO_S = SR * 2 + (SR / HOP) * 16   # 16 500 B / audio-second


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------ inputs

def synth_clips_gpu(n_clips, seed, device):
    """Synthetic corpus generated on the GPU with torch (plumbing, not the product):
    40 % tone, 30 % noise, 20 % chirp, 10 % composite; all passed through G.711 mu-law."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(20180610 + seed)
    out = torch.empty((n_clips, N_SAMP), dtype=torch.int16, device=device)
    t = torch.arange(N_SAMP, device=device, dtype=torch.float32) / SR
    chunk = 250
    for c0 in range(0, n_clips, chunk):
        n = min(chunk, n_clips - c0)
        u = torch.rand((n, 8), device=device, generator=g)
        kind = u[:, 0:1]
        f = 200.0 + u[:, 1:2] * 3200.0
        amp = 0.05 + u[:, 2:3] * 0.85
        tone = amp * torch.sin(2 * torch.pi * f * t + 2 * torch.pi * u[:, 3:4])
        sigma = 0.01 + u[:, 4:5] * 0.29
        noise = torch.clamp(torch.randn((n, N_SAMP), device=device, generator=g) * sigma, -1, 1)
        f0, f1 = 200.0, 0.85 * SR / 2
        chirp = (0.1 + 0.7 * u[:, 5:6]) * torch.sin(2 * torch.pi * (f0 * t + 0.5 * (f1 - f0) / SECONDS * t * t))
        comp = 0.5 * tone + 0.3 * chirp + 0.2 * noise
        x = torch.where(kind < 0.4, tone, torch.where(kind < 0.7, noise, torch.where(kind < 0.9, chirp, comp)))
        pcm = torch.round(torch.clamp(x, -1, 1) * 32767.0).to(torch.int32)
        # G.711 mu-law encode -> decode ("telephony ulaw-decoded")
        sign = pcm < 0
        mag = torch.clamp(pcm.abs(), max=32635) + 0x84
        exp = (torch.floor(torch.log2(mag.float())).to(torch.int32) - 7).clamp(0, 7)
        mant = (mag >> (exp + 3)) & 0x0F
        dec = (((mant << 3) + 0x84) << exp) - 0x84
        out[c0:c0 + n] = torch.where(sign, -dec, dec).to(torch.int16)
        del tone, noise, chirp, comp, x, pcm, mag, exp, mant, dec
    return out.reshape(-1)


def synth_clips_cpu(n_distinct, seed):
    from asterisk_tiresias_b200 import synth
    return np.stack([synth.make_clip(seed * 100003 + i, SECONDS, SR, ulaw=True) for i in range(n_distinct)])


# ------------------------------------------------------------------------------------ clocks

class ClockSampler:
    """nvidia-smi sampled while the timed region runs (B200_PROFILING.md clocks line)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.15:
                continue
            p = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(p[0])); mx = max(mx, float(p[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ reference arm

def run_reference(args):
    """The reference's CPU path (oracle port of libaubio) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import pyoracle as po
    n_thr = os.cpu_count() or 1
    plan = po.Plan(WIN, HOP, 40, 2, SR)
    pool = synth_clips_cpu(32, 1)
    # calibrate: one thread, 4 clips
    t = time.time(); plan.extract_batch(pool[:4].reshape(-1), np.arange(5, dtype=np.uint64) * N_SAMP, n_threads=1, want_y=False)
    per_clip = (time.time() - t) / 4
    target_s = 4.0
    n_clips = int(max(n_thr, min(10000, target_s / per_clip * n_thr)))
    pcm = pool[np.arange(n_clips) % pool.shape[0]].reshape(-1)
    off = np.arange(n_clips + 1, dtype=np.uint64) * N_SAMP
    for _ in range(args.warmup):
        plan.extract_batch(pcm, off, n_threads=n_thr, want_y=False)
    t0 = time.time()
    for _ in range(args.steps):
        plan.extract_batch(pcm, off, n_threads=n_thr, want_y=False)
    dt = (time.time() - t0) / args.steps
    value = n_clips * SECONDS / dt
    sample = f"{n_clips} of 10000 clips per step (32 distinct synthetic clips tiled), {n_thr} threads, oracle restatement of libaubio"
    line = {
        "impl": "reference", "metric": "audio_seconds_fingerprinted_per_second", "value": value, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(n_clips),
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": n_thr, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def workload_config(n_clips):
    return {"workload": "BASELINE config[1]: batch MFCC extraction of 10k synthetic 30 s 8 kHz mono PCM16 clips (G.711 mu-law decoded)",
            "clips_per_gpu": n_clips, "seconds_per_clip": SECONDS, "samplerate": SR, "win": WIN, "hop": HOP, "n_filters": 40,
            "n_coefs": 2, "l2": "inputs (4.8 GB per step) larger than L2"}


# ------------------------------------------------------------------------------------ match bench

def match_bench(ctx, args, rank, world, device, dist):
    """Secondary: match queries/s against a synthetic DB sharded by uuid over the ranks."""
    import torch
    from asterisk_tiresias_b200 import capi
    total_fps = args.match_fps if args.match_fps > 0 else 10_000_000     # BASELINE metric: 10M-fingerprint DB at every N
    F_db, Q, F_q = 94, args.match_queries, 94
    # uuids are random 128-bit values; shard s owns the uuids with tir_shard_of == s.  Generating the
    # shard directly (same distribution, n/world each) avoids materialising the whole DB on every rank.
    n_local = total_fps // world + (1 if rank < total_fps % world else 0)
    g = torch.Generator(device=device); g.manual_seed(991 + rank)
    rows = n_local * F_db
    uu = torch.randint(0, 256, (n_local, 16), dtype=torch.uint8, device=device, generator=g)
    # y1 ~ U(15.5, 18.5) (the ranges SURVEY.md 8a measured on 8 kHz material), in micro-units
    v1 = torch.randint(15_500_000, 18_500_000, (rows,), dtype=torch.int32, device=device, generator=g)
    v2 = torch.randint(-5_000_000, 20_000_000, (rows,), dtype=torch.int32, device=device, generator=g)
    row_off = (torch.arange(n_local + 1, device=device, dtype=torch.int64) * F_db)
    torch.cuda.synchronize()
    t0 = time.time()
    ctx.db_load_dev(n_local, uu.data_ptr(), row_off.data_ptr(), v1.data_ptr(), v2.data_ptr(), rows)
    build_s = time.time() - t0
    # queries (identical on every rank): 10 % exact copies of DB entries of rank 0, 10 % noisy copies, 80 % unrelated
    gq = torch.Generator(device=device); gq.manual_seed(4242)
    qv = torch.randint(15_500_000, 18_500_000, (Q, F_q), dtype=torch.int32, device=device, generator=gq).double() * 1e-6
    if world > 1:
        src = torch.zeros((Q // 5, F_q), dtype=torch.float64, device=device)
        if rank == 0:
            src = v1.view(n_local, F_db)[: Q // 5].double() * 1e-6
        dist.broadcast(src, 0)
    else:
        src = v1.view(n_local, F_db)[: Q // 5].double() * 1e-6
    qv[: Q // 10] = src[: Q // 10]
    qv[Q // 10: Q // 5] = src[Q // 10: Q // 5] + torch.randn((Q // 5 - Q // 10, F_q), device=device, generator=gq, dtype=torch.float64) * 3e-4
    # the engine takes mfcc coefficients; invert y = 10*log10|c| so that the device recomputes exactly these y
    q2 = torch.rand((Q, F_q), device=device, generator=gq, dtype=torch.float64) * 25.0 - 5.0   # max2 of every frame: all distinct
    coef = torch.stack([torch.pow(10.0, qv / 10.0).float(), torch.pow(10.0, q2 / 10.0).float()], dim=2).contiguous()
    qv = 10.0 * torch.log10(coef[:, :, 0].double())     # the y the device will recompute from the float coefficients
    foff = np.arange(Q + 1, dtype=np.uint64) * F_q
    d_hits = torch.zeros(Q * 24, dtype=torch.uint8, device=device)
    d_gather = torch.zeros(world * Q * 24, dtype=torch.uint8, device=device)
    d_final = torch.zeros(Q * 24, dtype=torch.uint8, device=device)

    # N > 1: the per-query winners (nq x 24 B per rank) are the only cross-GPU traffic.  Default: the
    # library's own exchange over NVLink peer memory (csrc/tir_p2p.cu: P2P stores into every peer's
    # gather buffer + flags, no collective call); --match-exchange nccl: all_gather + merge kernel.
    p2p = None
    exchange = "none (one GPU)"
    if world > 1:
        exchange = "nccl all_gather + tir_merge_hits_dev"
        if args.match_exchange == "p2p":
            try:
                p2p = capi.P2P(ctx, rank, world, Q)
                mine = torch.frombuffer(bytearray(p2p.handle()), dtype=torch.uint8).to(device)
                allh = torch.zeros(world * 64, dtype=torch.uint8, device=device)
                dist.all_gather_into_tensor(allh, mine)
                blob = allh.cpu().numpy().tobytes()
                p2p.connect([blob[64 * r: 64 * (r + 1)] for r in range(world)])
                ok = torch.ones(1, device=device)
            except Exception as e:   # e.g. no peer access between the devices: say so and use NCCL
                sys.stderr.write(f"[rank {rank}] tir_p2p unavailable ({e}); NCCL exchange\n")
                p2p, ok = None, torch.zeros(1, device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if float(ok.item()) < 1.0:
                p2p = None
            else:
                exchange = "tir_p2p (NVLink peer stores + flags, no collective call)"

    def step(coefs=1, nq=Q, use_p2p=True):
        if p2p is not None and use_p2p:
            p2p.match_dev(coef.data_ptr(), foff[: nq + 1], d_final.data_ptr(), coefs, 0.001)
            return
        ctx.match_dev(coef.data_ptr(), foff[: nq + 1], d_hits.data_ptr(), coefs, 0.001)
        if world > 1:
            dist.all_gather_into_tensor(d_gather[: world * nq * 24], d_hits[: nq * 24])
            ctx.merge_hits_dev(d_gather.data_ptr(), world, nq, d_final.data_ptr())
        # (one GPU: the shard's winners are the answer, nothing to merge)

    def timed(coefs, nq, steps, use_p2p=True):
        for _ in range(3):
            step(coefs, nq, use_p2p)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        l0 = ctx.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step(coefs, nq, use_p2p)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        k_ms = ctx.last_kernel_ms(1)
        launches = (ctx.launches - l0) / steps
        if world > 1:
            tt = torch.tensor([ms], device=device); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms = float(tt.item())
        return ms, k_ms, launches

    ctx.set_profiling(True)
    # per-query path (coefs = 2: every frame has its own max2 bounds), on a slice of the queries
    nq2 = max(1, min(Q, args.match_queries_coefs2))
    ms2, k_ms2, _ = timed(2, nq2, max(1, min(args.steps, 3)))
    # headline: coefs = 1, what the dialplan application passes (src/application_handler.c:180)
    nccl_ms = None
    if p2p is not None:   # the same batch through NCCL, for comparison (its winners are checked against the p2p ones)
        nccl_ms, _, _ = timed(1, Q, max(args.steps, 5), use_p2p=False)
        nccl_hits = d_final.clone()
    ms, kernel_ms, launches = timed(1, Q, max(args.steps, 5))
    if p2p is not None:
        same = bool(torch.equal(nccl_hits, d_final)) and p2p.error() == 0
        exchange += f"; identical to the NCCL exchange: {same}"
    hits = (d_final if world > 1 else d_hits).cpu().numpy().view(capi.HIT_DTYPE)
    search_e2e = search_bench(ctx, args, rank, world, device, dist, Q)
    cpu_match = match_cpu_baseline(args) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    # verification on rank 0 / single GPU: brute-force restatement of the vote for a few queries
    verified = None
    if world == 1:
        verified = verify_match(hits, qv.cpu().numpy(), v1.view(n_local, F_db), uu, n_check=4)
    # algorithmic bytes: F_q*8 + sum_k (16 + 8*R_k) + 24 per query; R_k from the data
    v1s = v1  # rows in window k of a query = rows with |v1 - k*1e6| <= 1000
    ks = torch.unique(torch.trunc(qv).to(torch.int64))
    R = {int(k): int(((v1s >= int(k) * 1_000_000 - 1000) & (v1s <= int(k) * 1_000_000 + 1000)).sum().item()) for k in ks}
    qk = torch.trunc(qv).to(torch.int64).cpu().numpy()
    alg = 0
    for q in range(Q):
        alg += F_q * 8 + 24 + sum(16 + 8 * R[int(k)] for k in np.unique(qk[q]))
    res = {"metric": "match_queries_per_second", "value": Q / (ms * 1e-3), "unit": "queries/s", "ms_per_batch": ms,
           "queries_per_batch": Q, "frames_per_query": F_q, "db_fingerprints_total": total_fps, "db_frames_per_fingerprint": F_db,
           "db_rows_this_rank": rows, "index_build_s": build_s, "coefs": 1, "tolerance": 0.001, "kernel_ms_rank0": kernel_ms,
           "launches_per_batch": launches, "found": int((hits["match_count"] > 0).sum()),
           "exchange": exchange, "nccl_exchange_ms_per_batch": nccl_ms,
           "path": "shared-window scan (distinct windows of the batch scanned once; DESIGN.md 4.3)",
           "per_query_path_coefs2": {"value": nq2 / (ms2 * 1e-3), "unit": "queries/s", "ms_per_batch": ms2, "queries_per_batch": nq2,
                                     "kernel_ms_rank0": k_ms2, "coefs": 2, "tolerance": 0.001},
           "self_matches_top": int((hits["match_count"][: Q // 10] > 0).sum()), "verified_queries": verified,
           "search_e2e": search_e2e, "cpu_baseline": cpu_match,
           "roofline": {"bound": "hbm", "achieved": alg / (kernel_ms * 1e-3) / 1e9 if kernel_ms and kernel_ms > 0 else None,
                        "unit": "GB/s", "algorithmic_bytes_per_batch_this_rank": alg,
                        "note": "SURVEY 8d charge F_q*8 + sum_k(16 + 8*R_k) + 24 per QUERY; the shared-window path reads each distinct window once per BATCH (6 B per row in it; the per-uuid patterns live in shared memory), so the charged figure can exceed the HBM peak"}}
    if p2p is not None:
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()      # no rank frees its region while a peer may still store into it
        p2p.close()
    del uu, v1, v2
    return res


def match_cpu_baseline(args):
    """The reference's match path on the host: its SQL text on the real SQLite (oracle, kind "reference
    SQL on libsqlite3"), one connection, one thread, on a BOUNDED database -- 5 000 fingerprints x 94
    frames (470 k rows; SQLite ingests ~30 k rows/s through the reference's textual INSERTs, so the
    10 M-fingerprint table of the GPU leg would take hours to build)."""
    from oracle import pyoracle as po
    from asterisk_tiresias_b200 import synth
    rng = np.random.default_rng(77)
    n_fp, F = 5000, 94
    sq = po.SqliteDB()
    t0 = time.time()
    for i in range(n_fp):
        y = np.stack([rng.uniform(15.5, 18.5, F), rng.uniform(-5.0, 20.0, F)], axis=1)
        sq.add_audio(synth.uuid_for(3_000_000 + i), y)
    ingest_s = time.time() - t0
    qs = [np.stack([rng.uniform(15.5, 18.5, F), rng.uniform(-5.0, 20.0, F)], axis=1) for _ in range(20)]
    t0 = time.time()
    found = sum(1 for y in qs if sq.search(y, 1, 0.001) is not None)
    dt = time.time() - t0
    return {"value": len(qs) / dt, "unit": "queries/s", "cores": 1, "kind": "reference SQL on libsqlite3 " + po.SqliteDB.sqlite_version(),
            "sample": f"{len(qs)} queries x {F} frames against {n_fp} fingerprints x {F} frames ({n_fp * F} rows), coefs 1, tolerance 0.001",
            "ingest_rows_per_s": n_fp * F / ingest_s, "found": found}


def search_bench(ctx, args, rank, world, device, dist, Q):
    """fp_search_fingerprint_info end to end through the C ABI (tir_search): Q query clips of 3 s
    (the dialplan default, src/application_handler.c:60) in pinned HOST memory -> extraction ->
    match against this rank's shard -> hits in host memory; with N ranks every rank searches its
    shard and the per-query winners are merged after an all-gather."""
    import ctypes as C
    import torch
    from asterisk_tiresias_b200 import capi, synth
    n = 3 * SR
    pool = np.stack([synth.make_clip(880000 + i, 3.0, SR, ulaw=True) for i in range(50)])
    h_pcm = torch.from_numpy(pool[np.arange(Q) % 50].reshape(-1).copy()).pin_memory()
    off = np.arange(Q + 1, dtype=np.uint64) * n
    h_hits = torch.zeros(Q * 24, dtype=torch.uint8).pin_memory()
    d_loc = torch.zeros(Q * 24, dtype=torch.uint8, device=device)
    d_all = torch.zeros(world * Q * 24, dtype=torch.uint8, device=device)
    d_out = torch.zeros(Q * 24, dtype=torch.uint8, device=device)
    L = capi.lib()

    def step():
        rc = L.tir_search(ctx._h, C.c_void_p(h_pcm.data_ptr()), off.ctypes.data_as(C.c_void_p), Q, 1, C.c_double(0.001), -1, -1,
                          C.c_void_p(h_hits.data_ptr()))
        if rc != 0:
            raise capi.TirError(rc, L.tir_last_error(ctx._h).decode())
        if world > 1:
            d_loc.copy_(h_hits, non_blocking=True)
            dist.all_gather_into_tensor(d_all, d_loc)
            ctx.merge_hits_dev(d_all.data_ptr(), world, Q, d_out.data_ptr())
            h_hits.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    steps = max(args.steps, 5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        tt = torch.tensor([ms], device=device); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms = float(tt.item())
    hits = h_hits.numpy().view(capi.HIT_DTYPE)
    return {"value": Q / (ms * 1e-3), "unit": "queries/s", "ms_per_batch": ms, "queries_per_batch": Q, "seconds_per_query_clip": 3.0,
            "h2d_bytes_per_step": int(h_pcm.numel() * 2), "d2h_bytes_per_step": Q * 24, "found": int((hits["match_count"] > 0).sum()),
            "api": "tir_search (host PCM16 -> winner uuid / match_count / frame_count in host memory)"}


def verify_match(hits, qv, v1, uu, n_check=4):
    """numpy/torch restatement of the vote for a few queries (the SQLite oracle cannot ingest a
    1M-fingerprint DB inside a bench run; it covers this path at small sizes in tests/)."""
    import torch
    from asterisk_tiresias_b200 import capi
    ok = 0
    uub = uu.cpu().numpy()
    order = np.lexsort(uub.T[::-1])
    rank_of = np.empty(len(order), np.int64); rank_of[order] = np.arange(len(order))
    rank_t = torch.from_numpy(rank_of).to(v1.device)
    for q in list(range(n_check // 2)) + list(range(len(qv) - n_check // 2, len(qv))):
        ks, w = np.unique(np.trunc(qv[q]).astype(np.int64), return_counts=True)
        votes = torch.zeros(v1.shape[0], dtype=torch.int64, device=v1.device)
        for k, wk in zip(ks, w):
            inwin = ((v1 >= int(k) * 1_000_000 - 1000) & (v1 <= int(k) * 1_000_000 + 1000)).any(dim=1)
            votes += inwin.to(torch.int64) * int(wk)
        best = int(votes.max().item())
        if best == 0:
            ok += int(hits["match_count"][q] == 0)
            continue
        cand = torch.nonzero(votes == best).flatten()
        win = int(cand[torch.argmax(rank_t[cand])].item())
        ok += int(hits["match_count"][q] == best and bytes(hits["uuid"][q].tolist()) == bytes(uub[win].tolist()))
    return {"checked": n_check, "identical": ok}


# ------------------------------------------------------------------------------------ our arm

_REAL_STDOUT = None


def claim_stdout():
    """The driver reads ONE JSON line from stdout; libraries (NCCL's version banner, ...) also write
    there from C.  Point fd 1 at stderr for the whole run and keep the real stdout for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=10000)
    ap.add_argument("--match-fps", type=int, default=0, help="fingerprints (uuids) in the match DB, total over ranks; 0 = 10M")
    ap.add_argument("--match-queries", type=int, default=1000)
    ap.add_argument("--match-exchange", choices=["p2p", "nccl"], default="p2p",
                    help="N > 1: how the per-query winners cross the GPUs (p2p = the library's NVLink peer-memory exchange)")
    ap.add_argument("--match-queries-coefs2", type=int, default=100, help="queries of the coefs=2 (per-query path) leg")
    ap.add_argument("--channels", type=int, default=1000, help="concurrent channel threads of the config[4] leg")
    ap.add_argument("--channels-db-fps", type=int, default=1_000_000)
    ap.add_argument("--no-match", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from asterisk_tiresias_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    # one explicit stream for everything that is timed: the library (a NULL handle would make it create
    # its own stream), torch's events and NCCL's collectives all run on it
    stream = torch.cuda.Stream(device=device)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = capi.Context(device=local_rank, win=WIN, hop=HOP, samplerate=SR, stream=stream.cuda_stream)
    ctx.set_profiling(True)

    n_clips = args.clips
    t0 = time.time()
    d_pcm = synth_clips_gpu(n_clips, rank, device)
    torch.cuda.synchronize()
    log(f"[rank {rank}] synthetic corpus: {n_clips} clips, {d_pcm.numel() * 2 / 1e9:.2f} GB, {time.time() - t0:.1f} s")
    off = np.arange(n_clips + 1, dtype=np.uint64) * N_SAMP
    F = n_clips * FRAMES_PER_CLIP
    d_coef = torch.empty((F, 2), dtype=torch.float32, device=device)
    d_vq = torch.empty((F, 2), dtype=torch.int32, device=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------
    for _ in range(args.warmup):
        ctx.extract_dev(d_pcm.data_ptr(), off, d_coef.data_ptr(), d_vq.data_ptr())
    barrier()
    sampler = ClockSampler(local_rank); sampler.start(); time.sleep(0.25)
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    w0 = time.time()
    e0.record()
    for _ in range(args.steps):
        ctx.extract_dev(d_pcm.data_ptr(), off, d_coef.data_ptr(), d_vq.data_ptr())
        kernel_ms.append(ctx.last_kernel_ms(0))       # waits for this step's kernel: events on the launching stream
    e1.record()
    barrier()
    w1 = time.time()
    ms_step = e0.elapsed_time(e1) / args.steps
    launches = int(ctx.launches - l0)
    # keep the GPU busy a little longer so that the 100 ms clock samples see it under load
    t_end = time.time() + 0.6
    while time.time() < t_end:
        ctx.extract_dev(d_pcm.data_ptr(), off, d_coef.data_ptr(), d_vq.data_ptr())
        torch.cuda.synchronize()
    clocks = sampler.stop(w0, time.time())
    if world > 1:
        tt = torch.tensor([ms_step], device=device); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms_step = float(tt.item())
    value = world * n_clips * SECONDS / (ms_step * 1e-3)
    k_ms = float(np.mean(kernel_ms))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = F * BYTES_PER_FRAME / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": 530.2 * F,   # ncu dram__bytes_read+write per launch, scaled from profiles/r1_extract_ncu_full_h.txt (529.6-530.2 B/frame)
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                "kernel": "tir_extract_kernel<512>", "kernel_ms": k_ms, "algorithmic_bytes_per_launch": F * BYTES_PER_FRAME,
                "note": "issue/latency-bound SIMT kernel on packed f32x2 instructions (float32 FFT reproduced operation for operation; FP32-pipe floor of the DAG = 24.7% of the HBM peak); see DESIGN.md 2.3 and profiles/"}

    # SURVEY 8d, config 2 "3 s variant": the same samples cut into ten times as many 3 s clips (94 frames
    # each, the last tile of every clip 30/32 full) -- exposes tile-tail and per-clip bookkeeping costs
    short_clips = None
    try:
        n3, s3 = n_clips * 10, N_SAMP // 10
        off3 = np.arange(n3 + 1, dtype=np.uint64) * s3
        F3 = n3 * ((s3 + HOP - 1) // HOP)
        d_coef3 = torch.empty((F3, 2), dtype=torch.float32, device=device)
        d_vq3 = torch.empty((F3, 2), dtype=torch.int32, device=device)
        for _ in range(6):   # the metadata of 10x as many clips outgrows every slot of the pinned staging ring once
            ctx.extract_dev(d_pcm.data_ptr(), off3, d_coef3.data_ptr(), d_vq3.data_ptr())
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            ctx.extract_dev(d_pcm.data_ptr(), off3, d_coef3.data_ptr(), d_vq3.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms3 = e0.elapsed_time(e1) / args.steps
        short_clips = {"clips_per_gpu": n3, "seconds_per_clip": SECONDS / 10, "frames_per_clip": int(F3 // n3),
                       "value_this_rank": n3 * (SECONDS / 10) / (ms3 * 1e-3), "unit": "audio-s/s", "ms_per_step": ms3}
        del d_coef3, d_vq3
    except Exception as e:   # a secondary figure must not take the headline down
        short_clips = {"error": str(e)}

    # parity is reported from the cpu_baseline leg below: the oracle's output for the clips it times
    # is compared with what the timed GPU run produced for the same clips (no other use of oracle/)
    parity = None
    g_coef_head = g_vq_head = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_s = min(n_clips, 500)
        g_coef_head = d_coef.view(n_clips, FRAMES_PER_CLIP, 2)[:n_s].cpu().numpy().reshape(-1, 2)
        g_vq_head = d_vq.view(n_clips, FRAMES_PER_CLIP, 2)[:n_s].cpu().numpy().reshape(-1, 2)

    # ---- end to end: host buffers through tir_extract ---------------------------------------
    e2e = None
    try:
        h_pcm = torch.empty(d_pcm.shape, dtype=torch.int16, pin_memory=True)
        h_pcm.copy_(d_pcm)
        h_coef = torch.empty((F, 2), dtype=torch.float32, pin_memory=True)
        h_vq = torch.empty((F, 2), dtype=torch.int32, pin_memory=True)
        torch.cuda.synchronize()
        import ctypes as C
        L = capi.lib()
        nf = C.c_uint64()
        offc = np.ascontiguousarray(off)

        def e2e_step():
            rc = L.tir_extract(ctx._h, C.c_void_p(h_pcm.data_ptr()), offc.ctypes.data_as(C.c_void_p), n_clips,
                               C.c_void_p(h_coef.data_ptr()), C.c_void_p(h_vq.data_ptr()), C.byref(nf))
            if rc != 0:
                raise capi.TirError(rc, L.tir_last_error(ctx._h).decode())
        for _ in range(2):
            e2e_step()
        barrier()
        e0.record()
        for _ in range(args.steps):
            e2e_step()
        e1.record()
        barrier()
        ms_e2e = e0.elapsed_time(e1) / args.steps
        if world > 1:
            tt = torch.tensor([ms_e2e], device=device); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms_e2e = float(tt.item())
        same = bool((h_vq.view(-1)[: 2 * FRAMES_PER_CLIP] == d_vq.view(-1)[: 2 * FRAMES_PER_CLIP].cpu()).all())
        e2e = {"value": world * n_clips * SECONDS / (ms_e2e * 1e-3), "unit": "audio-s/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": int(d_pcm.numel() * 2 + (n_clips + 1) * 20), "d2h_bytes_per_step": int(F * 16),
               "api": "tir_extract (host PCM16 in pinned memory -> coefficients + hashes in host memory)", "matches_device_run": same}
        # the same clips as the G.711 bytes they were decoded from (tir_extract_ulaw): half the H2D traffic
        try:
            sgn = d_pcm < 0
            mag = torch.clamp(d_pcm.to(torch.int32).abs(), max=32635) + 0x84
            ex_ = (torch.floor(torch.log2(mag.float())).to(torch.int32) - 7).clamp(0, 7)
            code = (~(torch.where(sgn, 0x80, 0) | (ex_ << 4) | ((mag >> (ex_ + 3)) & 0x0F)) & 0xFF).to(torch.uint8)
            h_law = torch.empty(code.shape, dtype=torch.uint8, pin_memory=True)
            h_law.copy_(code)
            del sgn, mag, ex_, code
            torch.cuda.synchronize()

            def ulaw_step():
                rc = L.tir_extract_ulaw(ctx._h, C.c_void_p(h_law.data_ptr()), offc.ctypes.data_as(C.c_void_p), n_clips,
                                        C.c_void_p(h_coef.data_ptr()), C.c_void_p(h_vq.data_ptr()), C.byref(nf))
                if rc != 0:
                    raise capi.TirError(rc, L.tir_last_error(ctx._h).decode())
            for _ in range(2):
                ulaw_step()
            barrier()
            e0.record()
            for _ in range(args.steps):
                ulaw_step()
            e1.record()
            barrier()
            ms_u = e0.elapsed_time(e1) / args.steps
            if world > 1:
                tt = torch.tensor([ms_u], device=device); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms_u = float(tt.item())
            e2e["ulaw_input"] = {"value": world * n_clips * SECONDS / (ms_u * 1e-3), "unit": "audio-s/s", "ms_per_step": ms_u,
                                 "h2d_bytes_per_step": int(h_law.numel() + (n_clips + 1) * 20), "d2h_bytes_per_step": int(F * 16),
                                 "api": "tir_extract_ulaw (the clips as the G.711 bytes they were decoded from; decoded on the device)",
                                 "matches_device_run": bool((h_vq.view(-1)[: 2 * FRAMES_PER_CLIP] == d_vq.view(-1)[: 2 * FRAMES_PER_CLIP].cpu()).all())}
            del h_law
        except Exception as ex:  # noqa: BLE001
            log("ulaw e2e leg failed:", ex)
        del h_pcm, h_coef, h_vq
    except Exception as ex:  # noqa: BLE001
        log("e2e leg failed:", ex)
        e2e = {"value": None, "unit": "audio-s/s", "error": str(ex)[:200], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}

    # ---- CPU baseline (rank 0, N=1 only): the oracle on a bounded sample, one thread --------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import pyoracle as po
        plan = po.Plan(WIN, HOP, 40, 2, SR)
        n_s = min(n_clips, 500)
        h = d_pcm.view(n_clips, N_SAMP)[:n_s].cpu().numpy().reshape(-1)
        t0 = time.time()
        oc, _, ov = plan.extract_batch(h, np.arange(n_s + 1, dtype=np.uint64) * N_SAMP, n_threads=1, want_y=False)
        dt = time.time() - t0
        # the checker side of the same leg: the three identity rates SURVEY.md H2 asks for (exact micro-unit
        # hash, the query side's trunc(max1), window membership at the default tolerance) + the MFCC error
        gc, gv = g_coef_head, g_vq_head
        k_g, k_o = np.trunc(gv[:, 0] / 1e6), np.trunc(ov[:, 0] / 1e6)
        in_g = np.abs(gv[:, 0] - np.rint(gv[:, 0] / 1e6) * 1e6) <= 1000
        in_o = np.abs(ov[:, 0] - np.rint(ov[:, 0] / 1e6) * 1e6) <= 1000
        relerr = np.abs(gc.astype(np.float64) - oc) / np.maximum(np.abs(oc), 1e-3 * np.abs(oc).max())
        parity = {"clips_checked": int(n_s), "frames_checked": int(gc.shape[0]),
                  "coef_bit_identical": float((gc.view(np.uint32) == oc.view(np.uint32)).mean()),
                  "mfcc_max_rel_err": float(relerr.max()), "hash_identical": float((gv == ov).mean()),
                  "hash_flips": int((gv != ov).sum()), "trunc_max1_identical": float((k_g == k_o).mean()),
                  "window_membership_identical": float((in_g == in_o).mean())}
        cpu_baseline = {"value": n_s * SECONDS / dt, "unit": "audio-s/s", "cores": 1, "kind": "port", "seconds": dt,
                        "host_cores_available": os.cpu_count(),
                        "sample": f"first {n_s} of the {n_clips} clips of this run, oracle restatement of libaubio pvoc+mfcc (float32 scalar C, -O2), 1 thread as in the reference"}

    # ---- match --------------------------------------------------------------------------------
    match = None
    del d_coef, d_vq, d_pcm
    torch.cuda.empty_cache()
    if not args.no_match:
        try:
            match = match_bench(ctx, args, rank, world, device, dist)
        except Exception as ex:  # noqa: BLE001
            log("match leg failed:", repr(ex))
            match = {"error": str(ex)[:300]}

    # ---- BASELINE config[4]: concurrent channels through the batcher (C++ client of the C ABI) ----
    concurrent = None
    tool = os.path.join(ROOT, "tools", "tir_concurrent_bench.bin")
    if rank == 0 and world == 1 and not args.no_match and os.path.exists(tool):
        try:
            torch.cuda.empty_cache()
            r = subprocess.run([tool, "--threads", str(args.channels), "--rounds", "20", "--wait-us", "1000", "--db-fps", str(args.channels_db_fps),
                                "--device", str(local_rank)], capture_output=True, text=True, timeout=600)
            concurrent = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": (r.stderr or r.stdout)[-300:]}
        except Exception as ex:  # noqa: BLE001
            concurrent = {"error": str(ex)[:300]}

    if rank == 0:
        line = {
            "metric": "audio_seconds_fingerprinted_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(n_clips), "clocks": clocks,
            "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity,
            "frames_per_second": world * F / (ms_step * 1e-3), "short_clips": short_clips, "match": match, "concurrent_channels": concurrent,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
