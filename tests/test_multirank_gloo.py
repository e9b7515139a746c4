"""World-size-2 test of the N>1 host logic on CPU (gloo): key broadcast + batch sharding."""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, "zig-tfhe_b200"))
    import tfhe_b200
    from tfhe_b200 import dist as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params = tfhe_b200.Params("tiny", 8, 1, 6, 2, 2)
    ck = sec = None
    if rank == 0:
        rng = np.random.default_rng(5)
        bshape, kshape = D.key_shapes(params)
        ck = tfhe_b200.CloudKey(rng.standard_normal(bshape), rng.integers(0, 2**32, kshape, dtype=np.uint32), 0x82080000)
        sec = rng.integers(0, 2, params.n + 1024, dtype=np.uint32)
    bsk, ksk, s = D.broadcast_cloud_key(params, ck, sec, torch.device("cpu"))
    h = hashlib.sha256(bsk.numpy().tobytes() + ksk.numpy().tobytes() + s.numpy().tobytes()).hexdigest()
    lo, hi = D.shard_range(1001, rank, world)
    out[rank] = (h, lo, hi)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_key_broadcast_and_sharding():
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, 29517 + os.getpid() % 1000, out), nprocs=world, join=True)
        (h0, lo0, hi0), (h1, lo1, hi1) = out[0], out[1]
    assert h0 == h1                                  # both ranks hold the same key material after one broadcast
    assert (lo0, hi0, lo1, hi1) == (0, 500, 500, 1001)


def test_shard_ranges_tile_the_batch():
    sys.path.insert(0, os.path.join(ROOT, "zig-tfhe_b200"))
    from tfhe_b200 import dist as D
    for total in (0, 1, 7, 65536, 1048576):
        for world in (1, 2, 4, 8):
            r = [D.shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
