"""Multi-GPU parity (needs >= 2 GPUs; skipped on a 1-GPU box): the slab-sharded
NCCL build must be BIT-IDENTICAL to the single-GPU build on every rank."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

KEYS = ("xyz", "colors", "opacity", "cov6", "sh", "weight", "is_parent")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from gaussiansplattingregistration_b200.hem import HemContext
    from gaussiansplattingregistration_b200.synth import synth
    c = synth(120_000, seed=61, density=5000.0, aspect=(2, 1, 0.5))
    L = 3
    single = HemContext(rank, seed=7)
    single.build(*c.arrays(), cluster_level=L)
    want = [single.level(l, True) for l in range(L + 1)]
    multi = HemContext(rank, seed=7)
    multi.set_option("shard_min_n", 20000)        # shard levels 0 and 1 of this small cloud, build level 2 whole
    multi.init_distributed()
    multi.build(*c.arrays(), cluster_level=L)
    ok = True
    msg = ""
    for l in range(L + 1):
        got = multi.level(l, True)
        for k in KEYS:
            if not np.array_equal(got[k], want[l][k], equal_nan=True):
                ok = False
                msg += f"rank {rank} level {l} {k} differs; "
    ret[rank] = (ok, msg, [multi.level_size(l) for l in range(L + 1)])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_bit_identical_to_single(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        ok, msg, sizes = ret[r]
        assert ok, msg
    assert all(ret[r][2] == ret[0][2] for r in range(world))
