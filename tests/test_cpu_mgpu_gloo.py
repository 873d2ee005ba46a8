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


class FakeSharer:
    """stands in for the library in the ShardedRunner protocol test: scripted payload sizes and
    hit counts, host memory instead of device memory"""

    def __init__(self, rank, world, script):
        import ctypes
        self.C = ctypes
        self.rank, self.world, self.script, self.k = rank, world, script, 0
        self.imported = []
        self.seen_payload = []
        self.truncated = False
        self.in_flight = False
        self.internal_cap = 5000
        self.redos = 0

    def setStream(self, s):
        pass

    def mgpuCollectTo(self, ptr, cap):
        total, _ = self.script[self.k]
        if total == 0:
            hdr = np.zeros(8, dtype=np.int64)
            hdr[0] = 0x47535331  # magic (the fake keeps the "nothing to run" status in its script)
            hdr[3] = 64
            self.C.memmove(ptr, hdr.ctypes.data, 64)
            return 64
        body = (np.arange(total, dtype=np.int64) * 7 + self.k) % 251
        buf = body.astype(np.uint8)
        hdr = np.zeros(8, dtype=np.int64)
        hdr[0] = 0x47535331
        hdr[3] = total
        buf[:64] = hdr.view(np.uint8)
        self.C.memmove(ptr, buf.ctypes.data, total)
        return total

    def _check(self, ptr, nbytes):
        total = self.script[self.k][0]
        got = np.ctypeslib.as_array((self.C.c_uint8 * nbytes).from_address(ptr)).copy()
        want = ((np.arange(total, dtype=np.int64) * 7 + self.k) % 251).astype(np.uint8)
        return bool(np.array_equal(got[64:nbytes], want[64:nbytes]) and got[:64].view(np.int64)[3] == total)

    def mgpuRunPayload(self, ptr, cap):  # rank 0: knows the whole batch
        total, _ = self.script[self.k]
        if total == 0:
            return -1
        self.seen_payload.append(self._check(ptr, total))
        self.in_flight = True
        return 0

    def mgpuEnqueuePayload(self, ptr, valid):  # receivers: blind enqueue of what arrived
        total, _ = self.script[self.k]
        if total == 0:
            return -1
        self.truncated = valid < total
        self.seen_payload.append(self._check(ptr, min(valid, total)))
        self.in_flight = True
        return 0

    def mgpuRedoPayload(self, ptr, total):
        assert self.truncated and total == self.script[self.k][0]
        self.seen_payload[-1] = self.seen_payload[-1] and self._check(ptr, total)
        self.truncated = False
        self.in_flight = True
        self.redos += 1

    def _hits(self):
        n = self.script[self.k][1] * (self.rank + 1)
        h = np.zeros(n, dtype=RAW_HIT_DTYPE)
        h["mask"] = np.arange(n) + 1
        h["solver"] = self.rank
        h["idx"] = self.k
        return h

    def mgpuEnqueueResult(self, ptr, cap):
        h = self._hits()
        if self.truncated:  # a truncated batch produces garbage: fewer hits than the real thing
            h = h[: len(h) // 2]
        over = len(h) > self.internal_cap  # the library's own buffers overflowed
        shown = h[: self.internal_cap] if over else h
        hdr = np.zeros(8, dtype=np.int64)
        hdr[0], hdr[1] = len(h), 1 if over else 0
        self.C.memmove(ptr, hdr.ctypes.data, 64)
        n = min(len(shown), cap)
        if n:
            self.C.memmove(ptr + 64, shown.ctypes.data, n * 16)
        return 64 + cap * 16

    def mgpuFinish(self):
        assert self.in_flight
        self.in_flight = False
        n = len(self._hits())
        if n > self.internal_cap:  # what finishRun does: grow the buffers and run again
            self.internal_cap = 2 * n
            return 1
        return 0

    def mgpuImport(self, hits):
        self.imported.append(np.array(hits, copy=True))

    def mgpuImportGathered(self, ptr, world, slot, counts):
        raw = np.ctypeslib.as_array((self.C.c_uint8 * (world * slot)).from_address(ptr))
        parts = [raw[r * slot + 64: r * slot + 64 + int(c) * 16] for r, c in enumerate(counts) if c]
        hits = np.concatenate(parts).view(RAW_HIT_DTYPE) if parts else np.zeros(0, dtype=RAW_HIT_DTYPE)
        self.imported.append(np.array(hits, copy=True))


def _runner_worker(rank, world, port, q):
    from gpusharesat_b200 import mgpu
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # (payload bytes, base hit count): growth beyond the prediction, shrink, nothing-to-run, hit overflow
    script = [(70000, 10), (400000, 3), (65600, 0), (0, 0), (300000, 9000), (66000, 5)]
    sh = FakeSharer(rank, world, script)
    runner = mgpu.ShardedRunner(sh, dist, rank, world, torch.device("cpu"), payload_cap=1 << 20, min_bcast=1 << 16)
    results = []
    for k in range(len(script)):
        sh.k = k
        results.append(runner.step())
    ok = all(sh.seen_payload) and len(sh.seen_payload) == 5
    ok = ok and results[3] is None
    if rank != 0:
        ok = ok and sh.redos >= 2  # batches 1 and 4 outgrew the predicted broadcast
    if rank == 0:
        steps = [k for k in range(len(script)) if script[k][0]]
        ok = ok and len(sh.imported) == len(steps)
        for imp, k in zip(sh.imported, steps):
            base = script[k][1]
            ok = ok and len(imp) == sum(base * (r + 1) for r in range(world))
            for r in range(world):
                part = imp[imp["solver"] == r]
                ok = ok and part["mask"].tolist() == list(range(1, base * (r + 1) + 1)) and bool(np.all(part["idx"] == k))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharded_runner_protocol_world2_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_runner_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}


class FakePeerSharer:
    """stands in for the library behind mgpu.PeerRunner: records what the runner hands it"""

    def __init__(self, rank):
        self.rank, self.blobs, self.steps = rank, None, 0

    def peerInit(self, rank, world, payload_cap, slot_hits):
        return bytes([rank]) * 16 + payload_cap.to_bytes(8, "little") + slot_hits.to_bytes(8, "little")

    def peerConnect(self, blobs):
        self.blobs = list(blobs)

    def peerEnqueue(self):
        self.steps += 1
        return -1 if self.steps == 2 else 0

    def peerFinish(self):
        return 100 * self.steps + self.rank

    def debugLastRunTimes(self):
        return [5.0, 1.0, 2.0, 30.0]


def _peer_runner_worker(rank, world, port, q):
    from gpusharesat_b200 import mgpu
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = FakePeerSharer(rank)
    runner = mgpu.PeerRunner(sh, dist, rank, world, payload_cap=1 << 20, slot_hits=777)
    # every rank got every rank's window handle, in rank order
    ok = sh.blobs == [bytes([r]) * 16 + (1 << 20).to_bytes(8, "little") + (777).to_bytes(8, "little") for r in range(world)]
    ok = ok and runner.step() == 100 + rank
    ok = ok and runner.step() is None  # "no clause yet": nothing is finished
    ok = ok and runner.step() == 300 + rank
    ok = ok and runner.device_us() == 25.0
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_peer_runner_bootstrap_world2_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_runner_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}
