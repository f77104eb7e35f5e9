"""Worker of tests/test_gpu_p2p.py::test_p2p_ipc_two_processes (one process per GPU, launched by
torch.distributed.run): the NVLink peer-memory exchange across PROCESSES (CUDA IPC handles exchanged through
an all_gather), checked on every rank against a context of its own that holds the whole table."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from asterisk_tiresias_b200 import capi, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pcm, off = synth.make_corpus(120, 2.0, first_index=52000, ragged=True)
    full = capi.Context(device=local)
    coef, vq = full.extract(pcm, off)
    fo = np.concatenate([[0], np.cumsum((np.diff(off.astype(np.int64)) + 255) // 256)]).astype(np.uint64)
    uu = np.stack([capi.uuid_to_bytes(synth.uuid_for(960000 + i)) for i in range(120)])
    full.db_load(uu, fo, vq[:, 0], vq[:, 1])
    mine = capi.Context(device=local)
    idx = np.array([a for a in range(120) if capi.shard_of(uu[a], world) == rank], np.int64)
    ro = np.zeros(idx.size + 1, np.uint64)
    rows = [np.arange(int(fo[a]), int(fo[a + 1])) for a in idx]
    ro[1:] = np.cumsum([x.size for x in rows])
    sel = np.concatenate(rows)
    mine.db_load(uu[idx], ro, vq[sel, 0], vq[sel, 1])
    q_clips = [pcm[int(off[i]):int(off[i + 1])] for i in range(0, 120, 4)] + [synth.make_clip(9100 + j, 1.1) for j in range(10)]
    Q = len(q_clips)
    qoff = np.zeros(Q + 1, np.uint64); qoff[1:] = np.cumsum([c.size for c in q_clips])
    qpcm = np.concatenate(q_clips)
    all_foff = np.concatenate([[0], np.cumsum([(c.size + 255) // 256 for c in q_clips])]).astype(np.uint64)
    p2p = capi.P2P(mine, rank, world, Q, max_frames=int(all_foff[-1]))
    h = torch.frombuffer(bytearray(p2p.handle()), dtype=torch.uint8).to(dev)
    allh = torch.zeros(world * 64, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allh, h)
    blob = allh.cpu().numpy().tobytes()
    p2p.connect([blob[64 * r: 64 * (r + 1)] for r in range(world)])
    per = (Q + world - 1) // world
    a, b = min(Q, rank * per), min(Q, (rank + 1) * per)
    for coefs, tol in ((1, 0.05), (2, 0.8), (1, 0.001), (1, 0.01)):
        ref = full.search(qpcm, qoff, coefs, tol)
        got = p2p.search(qpcm[int(qoff[a]):int(qoff[b])], qoff[a:b + 1] - qoff[a], a, all_foff, coefs, tol)
        assert p2p.error() == 0
        assert np.array_equal(got["match_count"], ref["match_count"]), (rank, coefs, tol)
        assert np.array_equal(got["frame_count"], ref["frame_count"])
        assert np.array_equal(got["uuid"][got["match_count"] > 0], ref["uuid"][ref["match_count"] > 0])
        # the match-only entry point on coefficients every rank already holds
        d_coef = torch.from_numpy(full.extract(qpcm, qoff)[0]).to(dev)
        d_fin = torch.zeros(Q * 24, dtype=torch.uint8, device=dev)
        p2p.match_dev(d_coef.data_ptr(), all_foff, d_fin.data_ptr(), coefs, tol)
        torch.cuda.synchronize()
        g2 = d_fin.cpu().numpy().view(capi.HIT_DTYPE)
        assert np.array_equal(g2["match_count"], ref["match_count"]) and np.array_equal(g2["uuid"][g2["match_count"] > 0], ref["uuid"][ref["match_count"] > 0])
    assert (ref["match_count"] > 0).sum() > 3
    torch.cuda.synchronize()
    dist.barrier()
    p2p.close()
    print("p2p-ipc-ok", rank, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
