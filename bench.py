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

def synth_clips_gpu(n_clips, seed, device, sr=SR):
    """Synthetic corpus generated on the GPU with torch (plumbing, not the product):
    40 % tone, 30 % noise, 20 % chirp, 10 % composite; all passed through G.711 mu-law."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(20180610 + seed)
    N_SAMP = int(sr * SECONDS)
    out = torch.empty((n_clips, N_SAMP), dtype=torch.int16, device=device)
    t = torch.arange(N_SAMP, device=device, dtype=torch.float32) / sr
    chunk = 250
    for c0 in range(0, n_clips, chunk):
        n = min(chunk, n_clips - c0)
        u = torch.rand((n, 8), device=device, generator=g)
        kind = u[:, 0:1]
        f = 200.0 + u[:, 1:2] * (3200.0 if sr <= 8000 else 6800.0)
        amp = 0.05 + u[:, 2:3] * 0.85
        tone = amp * torch.sin(2 * torch.pi * f * t + 2 * torch.pi * u[:, 3:4])
        sigma = 0.01 + u[:, 4:5] * 0.29
        noise = torch.clamp(torch.randn((n, N_SAMP), device=device, generator=g) * sigma, -1, 1)
        f0, f1 = 200.0, 0.85 * sr / 2
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
    # the stated config: every clip of the step (10 000 x 30 s; a few seconds per step on the box's cores)
    n_clips = args.clips
    pcm = pool[np.arange(n_clips) % pool.shape[0]].reshape(-1)
    off = np.arange(n_clips + 1, dtype=np.uint64) * N_SAMP
    for _ in range(args.warmup):
        plan.extract_batch(pcm, off, n_threads=n_thr, want_y=False)
    t0 = time.time()
    for _ in range(args.steps):
        plan.extract_batch(pcm, off, n_threads=n_thr, want_y=False)
    dt = (time.time() - t0) / args.steps
    value = n_clips * SECONDS / dt
    sample = f"all {n_clips} clips per step (32 distinct synthetic clips tiled), {n_thr} threads, oracle restatement of libaubio"
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

def _uuid_keys(uu):
    """[n,16] uint8 (cuda) -> two int64 keys whose signed order is the byte order of the uuids."""
    import torch
    w = uu.to(torch.int64)
    hi = torch.zeros(uu.shape[0], dtype=torch.int64, device=uu.device)
    lo = torch.zeros_like(hi)
    for i in range(8):
        hi = (hi << 8) | w[:, i]
        lo = (lo << 8) | w[:, 8 + i]
    flip = torch.tensor(-(2 ** 63), dtype=torch.int64, device=uu.device)
    return hi ^ flip, lo ^ flip


def _uuid_sort(uu):
    """ascending byte order of the shard's uuids (== SQLite's text order of the canonical form)"""
    import torch
    hi, lo = _uuid_keys(uu)
    o1 = torch.sort(lo, stable=True).indices
    o2 = torch.sort(hi[o1], stable=True).indices
    return o1[o2]


class BruteForce:
    """torch restatement of the vote (src/fp_handler.c:287-373) over this rank's shard, independent of the
    library's index and kernels: 'uuid has a row in window k' by comparing every stored row, votes =
    weights . bits, winner = greatest (count, uuid bytes).  Used only to verify the timed results."""

    def __init__(self, v1, v2, uu, F_db):
        self.v1, self.v2, self.F_db = v1.view(-1, F_db), v2.view(-1, F_db), F_db
        self.order = _uuid_sort(uu)
        self.uu_sorted = uu[self.order]
        self.n = uu.shape[0]

    def windows_bits(self, ks, T):
        """-> [K, n] float16 in sorted-uuid order: uuid has a row with |v1 - k*1e6| <= T"""
        import torch
        out = torch.zeros((len(ks), self.n), dtype=torch.float16, device=self.v1.device)
        step = max(1, (64 << 20) // self.F_db)
        for a0 in range(0, self.n, step):
            blk = self.v1[a0:a0 + step]
            for i, k in enumerate(ks):
                c = int(k) * 1_000_000
                out[i, a0:a0 + step] = ((blk >= c - T) & (blk <= c + T)).any(dim=1).to(torch.float16)
        return out[:, self.order]

    def winners_coefs1(self, y1, T):
        """y1 [Q, F] float64 (max1 of every query frame) -> (count int64 [Q], uuid uint8 [Q,16])"""
        import torch
        qk = torch.trunc(torch.where(torch.isfinite(y1), y1, torch.zeros_like(y1))).to(torch.int64)
        ks = torch.unique(qk)
        W = (qk[:, :, None] == ks[None, None, :]).sum(dim=1).to(torch.float16)          # [Q, K]
        B = self.windows_bits(ks.tolist(), T)                                            # [K, n]
        Q = y1.shape[0]
        best = torch.zeros(Q, dtype=torch.int64, device=y1.device)
        arg = torch.zeros(Q, dtype=torch.int64, device=y1.device)
        C = max(1, (256 << 20) // max(Q, 1))
        for c0 in range(0, self.n, C):
            votes = W @ B[:, c0:c0 + C]                                                  # exact: small integers in fp16
            m = votes.max(dim=1).values
            last = votes.shape[1] - 1 - torch.argmax((votes == m[:, None]).flip(1).to(torch.uint8), dim=1)
            mi = m.to(torch.int64)
            take = (mi >= best) & (mi > 0)        # later chunks hold greater uuids: ties go to them
            best = torch.where(take, mi, best)
            arg = torch.where(take, last + c0, arg)
        uuid = self.uu_sorted[arg] if self.n else torch.zeros((Q, 16), dtype=torch.uint8, device=y1.device)
        uuid = torch.where((best > 0)[:, None], uuid, torch.zeros_like(uuid))
        return best, uuid

    def winners_coefs2(self, y1, y2, T):
        """per-frame windows on both columns (coefs = 2), one vote per frame and uuid"""
        import torch
        dev = y1.device
        qk = torch.trunc(y1).to(torch.int64)
        ks = torch.unique(qk).tolist()
        rank_of = torch.empty(self.n, dtype=torch.int64, device=dev)
        rank_of[self.order] = torch.arange(self.n, device=dev)
        cand = []
        step = max(1, (64 << 20) // self.F_db)
        for a0 in range(0, self.n, step):
            blk = self.v1[a0:a0 + step]
            m = torch.zeros_like(blk, dtype=torch.bool)
            for k in ks:
                c = int(k) * 1_000_000
                m |= (blk >= c - T) & (blk <= c + T)
            idx = torch.nonzero(m)
            cand.append((rank_of[idx[:, 0] + a0], blk[m], self.v2[a0:a0 + step][m]))
        cu = torch.cat([c[0] for c in cand]); c1 = torch.cat([c[1] for c in cand]).to(torch.int64); c2 = torch.cat([c[2] for c in cand]).to(torch.int64)
        Q = y1.shape[0]
        best = torch.zeros(Q, dtype=torch.int64, device=dev)
        arg = torch.zeros(Q, dtype=torch.int64, device=dev)
        lo2 = torch.round((y2 - T * 1e-6) * 1e6).to(torch.int64)
        hi2 = torch.round((y2 + T * 1e-6) * 1e6).to(torch.int64)
        for q in range(Q):
            kc = qk[q] * 1_000_000
            m = (c1[None, :] >= (kc - T)[:, None]) & (c1[None, :] <= (kc + T)[:, None]) & \
                (c2[None, :] >= lo2[q][:, None]) & (c2[None, :] <= hi2[q][:, None])
            fj = torch.nonzero(m)
            if fj.shape[0] == 0:
                continue
            key = torch.unique(fj[:, 0] * self.n + cu[fj[:, 1]])                        # (frame, uuid) once
            u, cnt = torch.unique(key % self.n, return_counts=True)
            mx = cnt.max()
            best[q] = mx
            arg[q] = u[cnt == mx].max()
        uuid = self.uu_sorted[arg]
        uuid = torch.where((best > 0)[:, None], uuid, torch.zeros_like(uuid))
        return best, uuid


def _global_winners(best, uuid, world, dist):
    """fold the per-rank (count, uuid) candidates: greatest count, ties -> greatest uuid bytes"""
    import torch
    if world == 1:
        return best, uuid
    Q = best.shape[0]
    gb = torch.zeros((world, Q), dtype=torch.int64, device=best.device)
    gu = torch.zeros((world, Q, 16), dtype=torch.uint8, device=best.device)
    dist.all_gather_into_tensor(gb, best.contiguous())
    dist.all_gather_into_tensor(gu, uuid.contiguous())
    b, u = gb[0].clone(), gu[0].clone()
    for r in range(1, world):
        hi_a, lo_a = _uuid_keys(u)
        hi_b, lo_b = _uuid_keys(gu[r])
        greater = (hi_b > hi_a) | ((hi_b == hi_a) & (lo_b > lo_a))
        take = (gb[r] > b) | ((gb[r] == b) & (b > 0) & greater)
        b = torch.where(take, gb[r], b)
        u = torch.where(take[:, None], gu[r], u)
    return b, u


def _compare(d_hits_bytes, best, uuid, F_q=None):
    """the library's tir_hit array (device bytes) against the brute-force winners -> identical count"""
    import torch
    from asterisk_tiresias_b200 import capi
    h = d_hits_bytes.cpu().numpy().view(capi.HIT_DTYPE)
    b, u = best.cpu().numpy(), uuid.cpu().numpy()
    same = (h["match_count"] == b) & (h["uuid"] == u).all(axis=1)
    if F_q is not None:
        same &= h["frame_count"] == F_q
    return {"checked": int(b.shape[0]), "identical": int(same.sum())}


def match_bench(ctx, args, rank, world, device, dist):
    """Secondary metric: match queries/s against a synthetic DB sharded by uuid over the ranks.  Every timed
    result is compared, for EVERY query and at every N, with a torch brute force over the stored rows."""
    import torch
    from asterisk_tiresias_b200 import capi
    total_fps = args.match_fps if args.match_fps > 0 else 10_000_000     # BASELINE metric: 10M-fingerprint DB at every N
    Q, F_q = args.match_queries, 94
    qv_box = {}

    def make_db(F_db, seed):
        # uuids are random 128-bit values; shard s owns the uuids with tir_shard_of == s.  Generating the
        # shard directly (same distribution, n/world each) avoids materialising the whole DB on every rank.
        n_local = total_fps // world + (1 if rank < total_fps % world else 0)
        g = torch.Generator(device=device); g.manual_seed(seed + rank)
        rows = n_local * F_db
        uu = torch.randint(0, 256, (n_local, 16), dtype=torch.uint8, device=device, generator=g)
        # y1 ~ U(15.5, 18.5) (the ranges SURVEY.md 8a measured on 8 kHz material), in micro-units
        v1 = torch.randint(15_500_000, 18_500_000, (rows,), dtype=torch.int32, device=device, generator=g)
        v2 = torch.randint(-5_000_000, 20_000_000, (rows,), dtype=torch.int32, device=device, generator=g)
        row_off = (torch.arange(n_local + 1, device=device, dtype=torch.int64) * F_db)
        torch.cuda.synchronize()
        t0 = time.time()
        ctx.db_load_dev(n_local, uu.data_ptr(), row_off.data_ptr(), v1.data_ptr(), v2.data_ptr(), rows)
        torch.cuda.synchronize()
        return uu, v1, v2, n_local, rows, time.time() - t0

    F_db = 94
    uu, v1, v2, n_local, rows, build_s = make_db(F_db, 991)
    # queries (identical on every rank): 10 % exact copies of DB entries of rank 0, 10 % noisy copies, 80 % unrelated
    gq = torch.Generator(device=device); gq.manual_seed(4242)
    qv = torch.randint(15_500_000, 18_500_000, (Q, F_q), dtype=torch.int32, device=device, generator=gq).double() * 1e-6
    src = v1.view(n_local, F_db)[: Q // 5].double() * 1e-6
    if world > 1:
        if rank != 0:
            src = torch.zeros_like(src)
        dist.broadcast(src, 0)
    qv[: Q // 10] = src[: Q // 10]
    qv[Q // 10: Q // 5] = src[Q // 10: Q // 5] + torch.randn((Q // 5 - Q // 10, F_q), device=device, generator=gq, dtype=torch.float64) * 3e-4
    # the engine takes mfcc coefficients; invert y = 10*log10|c| so that the device recomputes exactly these y
    q2 = torch.rand((Q, F_q), device=device, generator=gq, dtype=torch.float64) * 25.0 - 5.0   # max2 of every frame: all distinct
    coef = torch.stack([torch.pow(10.0, qv / 10.0).float(), torch.pow(10.0, q2 / 10.0).float()], dim=2).contiguous()
    qv = 10.0 * torch.log10(coef[:, :, 0].double())     # the y the device will recompute from the float coefficients
    q2 = 10.0 * torch.log10(coef[:, :, 1].double())
    foff = np.arange(Q + 1, dtype=np.uint64) * F_q
    d_hits = torch.zeros(Q * 24, dtype=torch.uint8, device=device)
    d_gather = torch.zeros(world * Q * 24, dtype=torch.uint8, device=device)
    d_final = torch.zeros(Q * 24, dtype=torch.uint8, device=device)

    # N > 1: the per-query winners (nq x 24 B per rank) are the only cross-GPU traffic of the match.  Default: the
    # library's own exchange over NVLink peer memory (csrc/tir_p2p.cu: P2P stores into every peer's gather buffer
    # + flags, folded by the last CTA of the match chain, no collective call); --match-exchange nccl: all_gather +
    # merge kernel.
    p2p = None
    exchange = "none (one GPU)"
    if world > 1:
        exchange = "nccl all_gather + tir_merge_hits_dev"
        if args.match_exchange == "p2p":
            try:
                p2p = capi.P2P(ctx, rank, world, Q, max_frames=Q * F_q)
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
                exchange = "tir_p2p (NVLink peer stores + flags, folded in the last CTA of the match chain; no collective call)"

    def step(coefs=1, nq=Q, use_p2p=True, tol=0.001, cf=None, local_only=False):
        cf = coef if cf is None else cf
        if p2p is not None and use_p2p:
            p2p.match_dev(cf.data_ptr(), foff[: nq + 1], d_final.data_ptr(), coefs, tol)
            return
        ctx.match_dev(cf.data_ptr(), foff[: nq + 1], d_hits.data_ptr(), coefs, tol)
        if world > 1 and not local_only:
            dist.all_gather_into_tensor(d_gather[: world * nq * 24], d_hits[: nq * 24])
            ctx.merge_hits_dev(d_gather.data_ptr(), world, nq, d_final.data_ptr())
        # (one GPU: the shard's winners are the answer, nothing to merge)

    def timed(coefs, nq, steps, use_p2p=True, tol=0.001, cf=None, local_only=False):
        # warm-up: a steady caller's chain is captured into a CUDA graph once its key has been seen twice on each of the
        # four staging slots (4 plain calls, 4 captures), replays from then on
        for _ in range(12):
            step(coefs, nq, use_p2p, tol, cf, local_only)
        torch.cuda.synchronize()
        k_ms = ctx.last_kernel_ms(1)     # CUDA events recorded by the library around the chain of the last warm-up batch
        # the timed batches run without those events: two event records per batch are a measurable part of the host's
        # enqueue time of a 38 us chain (one graph launch)
        ctx.set_profiling(False)
        if world > 1:
            dist.barrier()
        l0 = ctx.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step(coefs, nq, use_p2p, tol, cf, local_only)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        ctx.set_profiling(True)
        launches = (ctx.launches - l0) / steps
        if world > 1:
            tt = torch.tensor([ms], device=device); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms = float(tt.item())
        return ms, k_ms, launches

    def result():
        return (d_final if world > 1 else d_hits)

    ctx.set_profiling(True)
    bf = BruteForce(v1, v2, uu, F_db)
    # per-query path (coefs = 2: every frame has its own max2 bounds), on a slice of the queries
    nq2 = max(1, min(Q, args.match_queries_coefs2))
    ms2, k_ms2, _ = timed(2, nq2, max(1, min(args.steps, 3)))
    b2, u2 = _global_winners(*bf.winners_coefs2(qv[:nq2], q2[:nq2], 1000), world, dist)
    verified2 = _compare(result()[: nq2 * 24], b2, u2, F_q)
    # headline: coefs = 1, what the dialplan application passes (src/application_handler.c:180)
    nccl_ms = None
    if p2p is not None:   # the same batch through NCCL, for comparison (its winners are checked against the p2p ones)
        nccl_ms, _, _ = timed(1, Q, max(args.steps, 5), use_p2p=False)
        nccl_hits = d_final.clone()
    ms, kernel_ms, launches = timed(1, Q, max(args.steps, 20))
    if p2p is not None:
        same = bool(torch.equal(nccl_hits, d_final)) and p2p.error() == 0
        exchange += f"; identical to the NCCL exchange: {same}"
    hits = result().cpu().numpy().view(capi.HIT_DTYPE)
    b1, u1 = _global_winners(*bf.winners_coefs1(qv, 1000), world, dist)
    verified = _compare(result(), b1, u1, F_q)
    # tolerance sweep: wider windows hold more rows -- the DB-bound regime, where the sharding pays
    sweep = {}
    for tol, T in ((0.01, 10_000), (0.05, 50_000)):
        ms_t, k_t, _ = timed(1, Q, max(args.steps, 20), tol=tol)
        bt, ut = _global_winners(*bf.winners_coefs1(qv, T), world, dist)
        sweep[str(tol)] = {"value": Q / (ms_t * 1e-3), "unit": "queries/s", "ms_per_batch": ms_t, "kernel_ms_rank0": k_t,
                           "verified_queries": _compare(result(), bt, ut, F_q)}
    # a batch of recordings of very different loudness: every query keeps its frames, shifted by a whole number of dB so that
    # the batch's frames take 64 distinct integer values of max1 -- 64 distinct windows, the most the shared-window path
    # holds (64-bit patterns); one window more and the batch takes the per-query kernel
    shift = (torch.arange(Q, device=device, dtype=torch.float64) % 16) * 4.0 - 35.0      # 16 shifts x 4 integers = 64 values
    shift[: Q // 5] = 0.0                                                                # the copies keep their level
    qv_w = qv + shift[:, None]
    coef_w = torch.stack([torch.pow(10.0, qv_w / 10.0).float(), coef[:, :, 1]], dim=2).contiguous()
    qv_w = 10.0 * torch.log10(coef_w[:, :, 0].double())
    n_win_w = int(torch.unique(torch.trunc(qv_w).to(torch.int64)).numel())
    ms_w, k_w, _ = timed(1, Q, max(args.steps, 20), cf=coef_w)
    bw, uw = _global_winners(*bf.winners_coefs1(qv_w, 1000), world, dist)
    many_windows = {"value": Q / (ms_w * 1e-3), "unit": "queries/s", "ms_per_batch": ms_w, "kernel_ms_rank0": k_w,
                    "distinct_windows": n_win_w, "verified_queries": _compare(result(), bw, uw, F_q),
                    "note": "shared-window path on 64-bit patterns (33..64 distinct windows); round 1: the per-query kernel from 33 windows on"}
    # honest roofline of the shared-window path: the bytes the algorithm must move per BATCH on this rank
    # (every distinct window's rows once: 2 B uid each, one 16 B range header per window and index block,
    # the coefficients in and the hits out) over the chain time
    ks = torch.unique(torch.trunc(qv).to(torch.int64)).tolist()
    R = [int(((v1 >= int(k) * 1_000_000 - 1000) & (v1 <= int(k) * 1_000_000 + 1000)).sum().item()) for k in ks]
    n_blocks = (n_local + 16383) // 16384
    alg = sum(R) * 2 + len(ks) * n_blocks * 16 + Q * F_q * 8 + Q * 24
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak))
    except OSError:
        pass
    achieved = alg / (ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "algorithmic_bytes_per_batch_this_rank": alg, "rows_in_windows_this_rank": sum(R), "distinct_windows": len(ks),
                "traffic": None, "traffic_note": "ncu dram bytes of the chain: see profiles/ (not measured in this run)",
                "note": "latency/launch-bound at this batch size: four dependent launches over a few MB; bytes = rows of the distinct windows x 2 B + range headers + coefficients + hits, per batch; time = the whole chain"}
    search_e2e = search_bench(ctx, args, rank, world, device, dist, Q, p2p, bf)
    cpu_match = match_cpu_baseline(args) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    res = {"metric": "match_queries_per_second", "value": Q / (ms * 1e-3), "unit": "queries/s", "ms_per_batch": ms,
           "queries_per_batch": Q, "frames_per_query": F_q, "db_fingerprints_total": total_fps, "db_frames_per_fingerprint": F_db,
           "db_rows_this_rank": rows, "index_build_s": build_s, "coefs": 1, "tolerance": 0.001, "kernel_ms_rank0": kernel_ms,
           "launches_per_batch": launches, "graph": (ctx.match_graph_stats() if world == 1 else None),
           "launch_note": "one GPU: the chain (offset copy, scratch clearing, 4 kernels) is replayed as ONE CUDA graph launch per batch; kernel_ms_rank0 brackets that launch",
           "found": int((hits["match_count"] > 0).sum()),
           "exchange": exchange, "nccl_exchange_ms_per_batch": nccl_ms,
           "path": "shared-window scan (distinct windows of the batch scanned once; DESIGN.md 4.3)",
           "verified_queries": verified, "verification": "torch brute force over every stored row of every rank, all queries (bench.py BruteForce)",
           "per_query_path_coefs2": {"value": nq2 / (ms2 * 1e-3), "unit": "queries/s", "ms_per_batch": ms2, "queries_per_batch": nq2,
                                     "kernel_ms_rank0": k_ms2, "coefs": 2, "tolerance": 0.001, "verified_queries": verified2},
           "tolerance_sweep": sweep, "many_distinct_windows": many_windows, "self_matches_top": int((hits["match_count"][: Q // 10] > 0).sum()),
           "search_e2e": search_e2e, "cpu_baseline": cpu_match, "roofline": roofline}
    del bf, uu, v1, v2
    torch.cuda.empty_cache()
    # The other way to use N GPUs when the table FITS one of them (10 M x 94 frames: 9.4 GB of master copy + 9.4 GB of
    # index): every rank holds the whole table and serves its own callers -- N independent replicas, no data-path
    # collective, no exchange ("replicas only"; the sharded legs above are for tables that do not fit, db_938_frames
    # below).  Every rank times its own 1 000-query batches against the same full table; value = N x Q / the slowest
    # rank's time; every rank's winners are verified against the brute force over the full table.
    res["replicated_table"] = None
    if world >= 2 and not args.no_db938:
        try:
            g = torch.Generator(device=device); g.manual_seed(991)
            rows_f = total_fps * F_db
            uu = torch.randint(0, 256, (total_fps, 16), dtype=torch.uint8, device=device, generator=g)
            v1 = torch.randint(15_500_000, 18_500_000, (rows_f,), dtype=torch.int32, device=device, generator=g)
            v2 = torch.randint(-5_000_000, 20_000_000, (rows_f,), dtype=torch.int32, device=device, generator=g)
            row_off = (torch.arange(total_fps + 1, device=device, dtype=torch.int64) * F_db)
            torch.cuda.synchronize()
            ctx.db_load_dev(total_fps, uu.data_ptr(), row_off.data_ptr(), v1.data_ptr(), v2.data_ptr(), rows_f)
            gr = torch.Generator(device=device); gr.manual_seed(777 + rank)              # every rank its own callers
            qr = torch.randint(15_500_000, 18_500_000, (Q, F_q), dtype=torch.int32, device=device, generator=gr).double() * 1e-6
            qr[: Q // 10] = v1.view(total_fps, F_db)[rank * Q: rank * Q + Q // 10].double() * 1e-6
            coef_r = torch.stack([torch.pow(10.0, qr / 10.0).float(), coef[:, :, 1]], dim=2).contiguous()
            qr = 10.0 * torch.log10(coef_r[:, :, 0].double())
            ms_r, k_r, _ = timed(1, Q, max(args.steps, 20), use_p2p=False, cf=coef_r, local_only=True)
            bfr = BruteForce(v1, v2, uu, F_db)
            br, ur = bfr.winners_coefs1(qr, 1000)
            vr = _compare(d_hits, br, ur, F_q)
            tt = torch.tensor([vr["checked"], vr["identical"]], device=device, dtype=torch.int64)
            dist.all_reduce(tt)
            res["replicated_table"] = {"value": world * Q / (ms_r * 1e-3), "unit": "queries/s", "ms_per_batch_slowest_rank": ms_r,
                                       "kernel_ms_rank0": k_r, "queries_per_batch_per_rank": Q, "collectives_in_data_path": 0,
                                       "verified_queries": {"checked": int(tt[0].item()), "identical": int(tt[1].item())},
                                       "note": "every rank holds the whole 10 M x 94 table (it fits) and serves its own batches; N replicas"}
            del bfr, uu, v1, v2
        except Exception as ex:  # noqa: BLE001
            res["replicated_table"] = {"error": repr(ex)[:300]}
        torch.cuda.empty_cache()
    # BASELINE config[3]/[4]'s table: 10 M fingerprints x 938 frames (30 s of audio each; 9.38 G rows, 94 GB of index)
    # -- only fits sharded (the index is built in passes: 18 B per row of master copy + index, 2 GB of scratch): N >= 2
    res["db_938_frames"] = None
    if world >= 2 and not args.no_db938:
        try:
            F2 = 938
            uu, v1, v2, n_local, rows2, build2 = make_db(F2, 1991)
            ms9, k9, _ = timed(1, Q, max(args.steps, 5))
            bf9 = BruteForce(v1, v2, uu, F2)
            b9, u9 = _global_winners(*bf9.winners_coefs1(qv, 1000), world, dist)
            v9 = _compare(result(), b9, u9, F_q)
            ms9b, _, _ = timed(1, Q, max(args.steps, 5), tol=0.01)
            b9b, u9b = _global_winners(*bf9.winners_coefs1(qv, 10_000), world, dist)
            res["db_938_frames"] = {"value": Q / (ms9 * 1e-3), "unit": "queries/s", "ms_per_batch": ms9, "kernel_ms_rank0": k9,
                                    "db_frames_per_fingerprint": F2, "db_rows_this_rank": rows2, "db_rows_total": total_fps * F2,
                                    "index_build_s": build2, "tolerance": 0.001, "verified_queries": v9,
                                    "tolerance_0.01": {"value": Q / (ms9b * 1e-3), "ms_per_batch": ms9b, "verified_queries": _compare(result(), b9b, u9b, F_q)}}
            del bf9, uu, v1, v2
        except Exception as ex:  # noqa: BLE001
            res["db_938_frames"] = {"error": repr(ex)[:300]}
        torch.cuda.empty_cache()
    if p2p is not None:
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()      # no rank frees its region while a peer may still store into it
        p2p.close()
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


def search_bench(ctx, args, rank, world, device, dist, Q, p2p, bf):
    """fp_search_fingerprint_info end to end through the C ABI: Q query clips of 3 s (the dialplan default,
    src/application_handler.c:60) in pinned HOST memory -> extraction -> match -> hits in host memory.
    One GPU: tir_search.  N ranks: tir_p2p_search -- every rank uploads and extracts only ITS Q/N clips, the
    coefficients cross the ranks inside the extraction kernel, every rank matches all Q against its shard and the
    winners are folded inside the match chain; every rank ends with all Q hits in host memory."""
    import ctypes as C
    import torch
    from asterisk_tiresias_b200 import capi, synth
    n = 3 * SR
    F_q = -(-n // HOP)
    pool = np.stack([synth.make_clip(880000 + i, 3.0, SR, ulaw=True) for i in range(50)])
    per = (Q + world - 1) // world
    a, b = min(Q, rank * per), min(Q, (rank + 1) * per)
    h_pcm = torch.from_numpy(pool[np.arange(a, b) % 50].reshape(-1).copy()).pin_memory()
    off = np.arange(b - a + 1, dtype=np.uint64) * n
    all_foff = np.arange(Q + 1, dtype=np.uint64) * F_q
    h_hits = torch.zeros(Q * 24, dtype=torch.uint8).pin_memory()
    L = capi.lib()
    fallback = world > 1 and p2p is None
    d_loc = d_all = d_out = h_all = off_all = None
    if fallback:   # no peer access: every rank searches all clips against its shard, NCCL all_gather + merge
        d_loc = torch.zeros(Q * 24, dtype=torch.uint8, device=device)
        d_all = torch.zeros(world * Q * 24, dtype=torch.uint8, device=device)
        d_out = torch.zeros(Q * 24, dtype=torch.uint8, device=device)
        h_all = torch.from_numpy(pool[np.arange(Q) % 50].reshape(-1).copy()).pin_memory()
        off_all = np.arange(Q + 1, dtype=np.uint64) * n

    def step():
        if world == 1:
            rc = L.tir_search(ctx._h, C.c_void_p(h_pcm.data_ptr()), off.ctypes.data_as(C.c_void_p), Q, 1, C.c_double(0.001), -1, -1,
                              C.c_void_p(h_hits.data_ptr()))
            if rc != 0:
                raise capi.TirError(rc, L.tir_last_error(ctx._h).decode())
        elif not fallback:
            p2p.search(None, off, a, all_foff, 1, 0.001, pcm_ptr=h_pcm.data_ptr(), hits_ptr=h_hits.data_ptr())
        else:
            rc = L.tir_search(ctx._h, C.c_void_p(h_all.data_ptr()), off_all.ctypes.data_as(C.c_void_p), Q, 1, C.c_double(0.001), -1, -1,
                              C.c_void_p(h_hits.data_ptr()))
            if rc != 0:
                raise capi.TirError(rc, L.tir_last_error(ctx._h).decode())
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
    # verification of every query: the brute force on the coefficients of the same clips (a separate extraction
    # through tir_extract on this rank; extraction parity itself is the parity object's business)
    pcm_all = pool[np.arange(Q) % 50].reshape(-1)
    cf, _ = ctx.extract(pcm_all, np.arange(Q + 1, dtype=np.uint64) * n)
    y1 = 10.0 * torch.log10(torch.from_numpy(cf[:, 0].astype(np.float64)).abs()).to(device).view(Q, F_q)
    bb, bu = _global_winners(*bf.winners_coefs1(y1, 1000), world, dist)
    verified = _compare(torch.from_numpy(h_hits.numpy().copy()), bb, bu, F_q)
    return {"value": Q / (ms * 1e-3), "unit": "queries/s", "ms_per_batch": ms, "queries_per_batch": Q, "seconds_per_query_clip": 3.0,
            "h2d_bytes_per_step_this_rank": int(h_pcm.numel() * 2) if not fallback else int(h_all.numel() * 2), "d2h_bytes_per_step": Q * 24,
            "found": int((hits["match_count"] > 0).sum()), "verified_queries": verified,
            "api": "tir_search (host PCM16 -> winner uuid / match_count / frame_count in host memory)" if world == 1 else
                   ("tir_p2p_search (each rank uploads + extracts its Q/N clips; coefficients and winners cross NVLink inside the kernels)" if not fallback
                    else "tir_search on every rank + NCCL all_gather + tir_merge_hits_dev (no peer access)")}


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
    ap.add_argument("--no-db938", action="store_true", help="skip the 10 M x 938-frame table leg (runs at N >= 2 only)")
    ap.add_argument("--no-wideband", action="store_true", help="skip the config[3] extraction leg (1024/512 at 16 kHz)")
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
    d_pcm = synth_clips_gpu(n_clips, 0, device)   # the same batch on every rank: the ranks' output checksums must agree
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
                "traffic": 529.2 * F,   # ncu dram__bytes_read+write per launch, scaled from the capture below (529.2 B/frame; round 1: 530.2)
                "traffic_source": "profiles/r2_extract_ncu_full_final.txt (ncu --set full capture of the same kernel, final tree of round 2; not measured in this run)",
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

    # BASELINE config[3]'s extraction: 16 kHz wideband audio with the larger window / hop (src/fp_handler.c:35-36,
    # the commented 1024 / 512 pair), half as many clips so that a step moves the same 4.8 GB
    wideband = None
    if not args.no_wideband:
        try:
            wsr, wwin, whop = 16000, 1024, 512
            wn = max(1, n_clips // 2)
            wctx = capi.Context(device=local_rank, win=wwin, hop=whop, samplerate=wsr, stream=stream.cuda_stream)
            wctx.set_profiling(True)
            w_pcm = synth_clips_gpu(wn, 7, device, sr=wsr)
            wsamp = int(wsr * SECONDS)
            woff = np.arange(wn + 1, dtype=np.uint64) * wsamp
            wF = wn * (-(-wsamp // whop))
            w_coef = torch.empty((wF, 2), dtype=torch.float32, device=device)
            w_vq = torch.empty((wF, 2), dtype=torch.int32, device=device)
            for _ in range(3):
                wctx.extract_dev(w_pcm.data_ptr(), woff, w_coef.data_ptr(), w_vq.data_ptr())
            torch.cuda.synchronize()
            wk = []
            e0.record()
            for _ in range(args.steps):
                wctx.extract_dev(w_pcm.data_ptr(), woff, w_coef.data_ptr(), w_vq.data_ptr())
                wk.append(wctx.last_kernel_ms(0))
            e1.record()
            torch.cuda.synchronize()
            wms = e0.elapsed_time(e1) / args.steps
            wbytes = wF * (whop * 2 + 16)
            wach = wbytes / (float(np.mean(wk)) * 1e-3) / 1e9
            wideband = {"workload": "BASELINE config[3] extraction: 16 kHz wideband, win 1024 / hop 512", "clips_per_gpu": wn, "seconds_per_clip": SECONDS,
                        "samplerate": wsr, "win": wwin, "hop": whop, "value_this_rank": wn * SECONDS / (wms * 1e-3), "unit": "audio-s/s",
                        "ms_per_step": wms, "frames_per_second": wF / (wms * 1e-3),
                        "roofline": {"bound": "hbm", "achieved": wach, "peak": peak, "unit": "GB/s", "frac": wach / peak, "kernel": "tir_extract_kernel<1024>",
                                     "kernel_ms": float(np.mean(wk)), "algorithmic_bytes_per_launch": wbytes, "traffic": None}}
            del w_pcm, w_coef, w_vq
            wctx.close()
            torch.cuda.empty_cache()
        except Exception as e:   # a secondary figure must not take the headline down
            wideband = {"error": repr(e)[:300]}

    # parity is reported from the cpu_baseline leg below: the oracle's output for the clips it times
    # is compared with what the timed GPU run produced for the same clips (no other use of oracle/)
    parity = None
    g_coef_head = g_vq_head = None
    # every rank ran the same batch: one 64-bit checksum of all its hashes and coefficient bit patterns; the ranks
    # must agree with rank 0, whose first 500 clips are compared with the oracle below
    ck = (d_vq.view(-1).to(torch.int64) * 1_000_003 + d_coef.view(torch.int32).view(-1).to(torch.int64)).sum()
    ck_all = [int(ck.item())]
    if world > 1:
        gl = torch.zeros(world, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(gl, ck.view(1))
        ck_all = [int(x) for x in gl.tolist()]
    if rank == 0 and not args.no_cpu_baseline:
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
    if rank == 0 and not args.no_cpu_baseline:   # (the oracle run yields the parity object at every N; it is REPORTED as cpu_baseline at N = 1 only)
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
                  "window_membership_identical": float((in_g == in_o).mean()),
                  "ranks_output_checksums_identical": len(set(ck_all)) == 1, "ranks_checked": len(ck_all),
                  "checksum": "sum over all frames of vq*1000003 + coefficient bits (int64), whole batch, every rank"}
        cpu_baseline = None if world > 1 else {"value": n_s * SECONDS / dt, "unit": "audio-s/s", "cores": 1, "kind": "port", "seconds": dt,
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
    # one GPU: tir_search_one on one context (+ the streaming variant: 20 ms chunks through tir_stream_*);
    # N GPUs: ONE process drives all N devices through tir_group_* (an Asterisk module is one process) against the
    # 10 M-fingerprint table -- rank 0 runs it while the other ranks, their work done, wait at the barrier
    concurrent = None
    tool = os.path.join(ROOT, "tools", "tir_concurrent_bench.bin")
    ctx.close()
    torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
    if rank == 0 and not args.no_match and os.path.exists(tool):
        def run_tool(extra):
            r = subprocess.run([tool, "--threads", str(args.channels), "--rounds", "20", "--wait-us", "1000", *extra], capture_output=True, text=True, timeout=900)
            return json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": (r.stderr or r.stdout)[-300:]}
        try:
            if world == 1:
                concurrent = run_tool(["--db-fps", str(args.channels_db_fps), "--device", str(local_rank)])
                concurrent["streaming"] = run_tool(["--db-fps", str(args.channels_db_fps), "--device", str(local_rank), "--stream", "1"])
            else:
                concurrent = run_tool(["--db-fps", str(args.match_fps if args.match_fps > 0 else 10_000_000), "--devices", str(world)])
        except Exception as ex:  # noqa: BLE001
            concurrent = {"error": str(ex)[:300]}

    if rank == 0:
        line = {
            "metric": "audio_seconds_fingerprinted_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(n_clips), "clocks": clocks,
            "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity,
            "frames_per_second": world * F / (ms_step * 1e-3), "short_clips": short_clips, "wideband_config3": wideband, "match": match,
            "concurrent_channels": concurrent,
        }
        # last in the line (the driver keeps the tail): the second half of the BASELINE metric, per N, with its verification
        if isinstance(match, dict) and "value" in match:
            v = lambda d: None if not d else f'{d["identical"]}/{d["checked"]}'   # noqa: E731
            line["match_scaling"] = {
                "n_gpus": world, "db": f'{match["db_fingerprints_total"]} fingerprints x {match["db_frames_per_fingerprint"]} frames',
                "coefs1_qps": match["value"], "coefs1_ms_per_batch": match["ms_per_batch"], "coefs1_verified": v(match["verified_queries"]),
                "coefs2_qps": match["per_query_path_coefs2"]["value"], "coefs2_verified": v(match["per_query_path_coefs2"]["verified_queries"]),
                "tol_0.01_qps": match["tolerance_sweep"]["0.01"]["value"], "tol_0.01_verified": v(match["tolerance_sweep"]["0.01"]["verified_queries"]),
                "tol_0.05_qps": match["tolerance_sweep"]["0.05"]["value"], "tol_0.05_verified": v(match["tolerance_sweep"]["0.05"]["verified_queries"]),
                "win64_qps": match["many_distinct_windows"]["value"], "win64_verified": v(match["many_distinct_windows"]["verified_queries"]),
                "search_e2e_qps": match["search_e2e"]["value"], "search_e2e_verified": v(match["search_e2e"]["verified_queries"]),
                "replicated_qps": (match.get("replicated_table") or {}).get("value"), "replicated_verified": v((match.get("replicated_table") or {}).get("verified_queries")),
                "db938_qps": (match.get("db_938_frames") or {}).get("value"), "db938_verified": v((match.get("db_938_frames") or {}).get("verified_queries")),
            }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
