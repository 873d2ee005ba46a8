"""CPU tier, world_size 2 over gloo: the multi-GPU exchange (payload broadcast from rank 0, hit
gather to rank 0) that bench.py --gpus N runs over NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gpusharesat_b200.api import RAW_HIT_DTYPE


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import ctypes as C
    from gpusharesat_b200 import mgpu
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = torch.device("cpu")
    rng = np.random.default_rng(5)
    params = rng.integers(0, 255, size=3 * 288, dtype=np.uint8)
    updates = rng.integers(0, 255, size=1000 * 12, dtype=np.uint8)
    collected = None
    if rank == 0:
        collected = (1, params.ctypes.data, params.size, updates.ctypes.data, 1000)
    status, p, u, n = mgpu.broadcast_batch(dist, rank, dev, collected)
    ok = status == 1 and n == 1000 and np.array_equal(p.numpy(), params) and np.array_equal(u.numpy()[: 12000], updates)
    # nothing-to-run batch
    status2, _, _, _ = mgpu.broadcast_batch(dist, rank, dev, None)
    ok = ok and status2 == -1
    # empty update list
    status3, p3, _, n3 = mgpu.broadcast_batch(dist, rank, dev, (0, params.ctypes.data, params.size, 0, 0) if rank == 0 else None)
    ok = ok and status3 == 0 and n3 == 0 and np.array_equal(p3.numpy(), params)
    # ragged hit lists: rank r contributes 10 * r + 3 hits
    mine = np.zeros(10 * rank + 3, dtype=RAW_HIT_DTYPE)
    mine["mask"] = np.arange(len(mine)) + 1
    mine["solver"] = rank
    mine["idx"] = np.arange(len(mine)) * world + rank
    got = mgpu.gather_hits(dist, rank, world, dev, mine)
    if rank == 0:
        ok = ok and len(got) == sum(10 * r + 3 for r in range(world))
        for r in range(world):
            part = got[got["solver"] == r]
            ok = ok and part["mask"].tolist() == list(range(1, 10 * r + 4))
    else:
        ok = ok and got is None
    # all-empty gather
    e = mgpu.gather_hits(dist, rank, world, dev, np.zeros(0, dtype=RAW_HIT_DTYPE))
    ok = ok and ((rank == 0 and len(e) == 0) or (rank != 0 and e is None))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_payload_broadcast_and_hit_gather_world2_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}
