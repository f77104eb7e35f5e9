"""world_size-2 gloo test of the sharded match plumbing on CPU: shard assignment, all-gather of the
per-shard tir_hit records, and the fold rule (greatest count, ties -> greatest uuid bytes).  The
per-shard engine here is the SQLite oracle (no GPU in this container); on the GPU box the same
plumbing runs over NCCL in bench.py and the fold is tir_merge_hits_dev (tests/test_gpu_match.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from asterisk_tiresias_b200 import capi, sharding, synth, synth_db


def fold_hits(gathered: np.ndarray) -> np.ndarray:
    """numpy statement of tir_merge_hits_dev: gathered [S, Q] of HIT_DTYPE -> [Q]."""
    out = gathered[0].copy()
    for s in range(1, gathered.shape[0]):
        for q in range(gathered.shape[1]):
            h, b = gathered[s, q], out[q]
            better = h["match_count"] > b["match_count"] or (
                h["match_count"] == b["match_count"] and h["match_count"] > 0 and bytes(h["uuid"]) > bytes(b["uuid"]))
            if better:
                out[q] = h
    return out


def _worker(rank, world, port, db, queries, params, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as po
    uu = np.stack([capi.uuid_to_bytes(u) for u, _ in db])
    owner = sharding.partition_by_uuid(uu, world)
    sq = po.SqliteDB()
    for (u, y), o in zip(db, owner):
        if o == rank:
            sq.add_audio(u, y)
    hits = np.zeros(len(queries), capi.HIT_DTYPE)
    for i, y in enumerate(queries):
        r = sq.search(y, *params, has_y=np.isfinite(y))
        hits[i]["frame_count"] = y.shape[0]
        if r is not None:
            hits[i]["uuid"] = capi.uuid_to_bytes(r["uuid"])
            hits[i]["match_count"] = r["match_count"]
    mine = torch.from_numpy(hits.view(np.uint8).copy())
    gathered = sharding.all_gather_hits(mine, world, dist).numpy().view(capi.HIT_DTYPE).reshape(world, len(queries))
    merged = fold_hits(gathered)
    if rank == 0:
        ret["merged"] = merged
        ret["owner_counts"] = np.bincount(owner, minlength=world)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("params", [(1, 0.01, -1, -1), (2, 0.6, -1, -1)])
def test_two_rank_sharded_match_equals_unsharded(oracle, params):
    rng = np.random.default_rng(31)
    db = synth_db.make_db(400, 5, 25, seed=31)
    for i in range(40):
        db.append((synth.uuid_for(6_000_000 + i), db[i][1].copy()))      # ties that straddle the shards
    queries = [db[int(rng.integers(0, len(db)))][1] for _ in range(8)] + [synth_db.random_y(rng, 30) for _ in range(6)]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, db, queries, params, ret), nprocs=2, join=True)
    merged = ret["merged"]
    assert ret["owner_counts"].sum() == len(db) and ret["owner_counts"].min() > len(db) // 4
    sq = oracle.SqliteDB()
    for u, y in db:
        sq.add_audio(u, y)
    for y, h in zip(queries, merged):
        exp = sq.search(y, *params, has_y=np.isfinite(y))
        if exp is None:
            assert h["match_count"] == 0
        else:
            assert (capi.bytes_to_uuid(h["uuid"]), int(h["match_count"]), int(h["frame_count"])) == \
                (exp["uuid"], exp["match_count"], exp["frame_count"])
