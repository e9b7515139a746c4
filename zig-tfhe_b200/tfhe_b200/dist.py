"""One-process-per-GPU plumbing (torch.distributed): key replication and batch sharding.

The hot path has no collective: bootstraps are independent given read-only keys, so the
only communication is ONE broadcast of the cloud key at load time (NCCL over NVLink on GPUs, gloo on CPU in the
tests).  This stands in for the reference's data-parallel thread pool (src/parallel/thread_pool.zig:39-83), which
shares the key through process memory instead.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import CloudKey, Params


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """contiguous slice of a batch owned by `rank` (same rule as the C library's in-process device split)"""
    return total * rank // world, total * (rank + 1) // world


def key_shapes(params: Params):
    return (params.n, 2 * params.L, 2, 1024), (1024 * params.iks_t * (1 << params.basebit), params.n + 1)


def broadcast_cloud_key(params: Params, ck: CloudKey | None, secret: np.ndarray | None, device: torch.device, src: int = 0):
    """Rank `src` passes its CloudKey (and optionally the concatenated secret key for test/bench decryption);
    every rank returns (bsk f64 tensor, ksk i32 tensor, secret i32 tensor | None) on `device`."""
    bsk_shape, ksk_shape = key_shapes(params)
    rank = dist.get_rank() if dist.is_initialized() else 0
    bsk = torch.empty(bsk_shape, dtype=torch.float64, device=device)
    ksk = torch.empty(ksk_shape, dtype=torch.int32, device=device)
    sec = torch.empty(params.n + 1024, dtype=torch.int32, device=device)
    has_secret = torch.zeros(1, dtype=torch.int32, device=device)
    if rank == src:
        bsk.copy_(torch.from_numpy(np.ascontiguousarray(ck.bootstrapping_key, dtype=np.float64).reshape(bsk_shape)))
        ksk.copy_(torch.from_numpy(np.ascontiguousarray(ck.key_switching_key, dtype=np.uint32).view(np.int32).reshape(ksk_shape)))
        if secret is not None:
            sec.copy_(torch.from_numpy(np.ascontiguousarray(secret, dtype=np.uint32).view(np.int32)))
            has_secret += 1
    if dist.is_initialized() and dist.get_world_size() > 1:
        for t in (bsk, ksk, sec, has_secret):
            dist.broadcast(t, src)
    return bsk, ksk, (sec if int(has_secret.item()) else None)
