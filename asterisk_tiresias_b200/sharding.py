"""Host-side plumbing of the multi-GPU match (one process per GPU): which shard owns a uuid and
how the per-shard winners of a query batch are exchanged.  The exchange is one all-gather of
n_queries x 24 bytes (tir_hit) per batch; the fold itself is tir_merge_hits_dev on the GPU."""
from __future__ import annotations

import numpy as np

from . import capi


def partition_by_uuid(uuids16: np.ndarray, n_shards: int) -> np.ndarray:
    """shard index of every uuid ([n,16] uint8) -- tir_shard_of (FNV-1a of the bytes mod n)."""
    return np.array([capi.shard_of(u, n_shards) for u in np.asarray(uuids16, dtype=np.uint8).reshape(-1, 16)], dtype=np.int64)


def all_gather_hits(hits_u8, world_size, dist):
    """hits_u8: torch uint8 tensor [n_queries*24] on this rank -> [world_size*n_queries*24]."""
    import torch
    out = torch.empty(world_size * hits_u8.numel(), dtype=torch.uint8, device=hits_u8.device)
    dist.all_gather_into_tensor(out, hits_u8)
    return out
