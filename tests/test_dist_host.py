"""CPU tests of the multi-GPU host logic with world_size 2 over gloo: the slab
partition every rank derives must tile the parents exactly and agree across
ranks, and the NCCL-id bootstrap must deliver rank 0's bytes to every rank."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gaussiansplattingregistration_b200 import _capi
    from gaussiansplattingregistration_b200.hem import exchange_unique_id, slab_range
    ok = True
    for n_parents in (0, 1, 2, 3, 7, 333_333, 1_000_822):
        b, e = slab_range(n_parents, rank, world)
        mine = torch.tensor([b, e], dtype=torch.int64)
        allr = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allr, mine)
        bounds = torch.stack(allr)
        ok &= int(bounds[0, 0]) == 0 and int(bounds[-1, 1]) == n_parents
        ok &= bool((bounds[1:, 0] == bounds[:-1, 1]).all())                   # contiguous, no gap / overlap
        sizes = bounds[:, 1] - bounds[:, 0]
        ok &= int(sizes.max()) == -(-n_parents // world)                       # equal chunks of ceil(np/G)
    fake = bytes(np.random.default_rng(5).integers(0, 256, _capi.NCCL_ID_BYTES, dtype=np.uint8))
    got, r, w = exchange_unique_id(make_id=lambda: fake)
    ok &= got == fake and r == rank and w == world
    ret[rank] = ok
    dist.destroy_process_group()


def test_slab_partition_and_id_bootstrap_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)) and len(ret) == world


def test_slab_range_rejects_bad_arguments():
    import pytest
    from gaussiansplattingregistration_b200._capi import HemError
    from gaussiansplattingregistration_b200.hem import slab_range
    assert slab_range(10, 0, 1) == (0, 10)
    assert [slab_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    with pytest.raises(HemError):
        slab_range(10, 4, 4)
    with pytest.raises(HemError):
        slab_range(-1, 0, 2)
