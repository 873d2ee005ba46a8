"""Multi-GPU exchange over peer memory (csrc/peer.cu, gss_peer_*).  The GPU box of the test tier has
ONE device, so the world-size-2 cases run two PROCESSES on that device: CUDA IPC, the stream memory
operations and the whole protocol (mailbox, peer loads of the batch, peer stores of the hits, done
flags, overflow second round) are exactly what runs across NVLink, only the wire is missing.
Rank 0's hand-over must be identical to an unsharded sharer fed the same calls."""
import multiprocessing as mp
import os
import traceback

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _build(sh, n, nsolvers):
    from gpusharesat_b200 import mkLit
    sh.setVarCount(n)
    sh.setCpuSolverCount(nsolvers)
    for i in range(n):
        if i % 7 == 0:
            sh.addClause(-1, [mkLit(i), mkLit((i + 1) % n, True), mkLit((i + 5) % n)])
        elif i % 3:
            sh.addClause(-1, [mkLit(i), mkLit((i + 1) % n, True)])
        else:
            sh.addClause(-1, [mkLit(i)])


def _push(shs, step, n, nsolvers):
    """same assignment traffic into every front-end in `shs`"""
    from gpusharesat_b200 import mkLit
    rng = np.random.default_rng(1000 + step)
    for s in range(nsolvers):
        for _ in range(int(rng.integers(1, 4))):
            vs = rng.choice(n, size=max(1, n // (s + 2)), replace=False)
            x = rng.random(len(vs))
            unset = [mkLit(int(v)) for v, xx in zip(vs, x) if xx < 0.1]
            sets = [mkLit(int(v), bool(xx < 0.8)) for v, xx in zip(vs, x) if xx >= 0.1]
            for sh in shs:
                sh.unsetSolverValues(s, unset)
                assert sh.trySetSolverValues(s, sets)
                assert sh.trySendAssignment(s) >= 0


def _pop_all(sh, nsolvers):
    out = []
    for s in range(nsolvers):
        while (x := sh.popReportedClause(s)) is not None:
            out.append((s, x[1], tuple(x[0])))
    return out


def _worker(rank, world, port, n, nsolvers, steps, wait_mode, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        os.environ["GPUSHARE_DEVICE"] = "0"
        os.environ["GPUSHARE_PEER_TIMEOUT_S"] = "30"
        if wait_mode == "kernel":
            os.environ["GPUSHARE_PEER_WAIT"] = "kernel"
        import torch.distributed as dist
        from gpusharesat_b200 import GpuClauseSharer, GpuClauseSharerOptions, mgpu
        dist.init_process_group("gloo", rank=rank, world_size=world)
        opts = dict(minGpuLatencyMicros=0, verbosity=0)
        sh = GpuClauseSharer(GpuClauseSharerOptions(**opts))
        sh.setShard(rank, world)
        _build(sh, n, nsolvers)
        single = None
        if rank == 0:
            single = GpuClauseSharer(GpuClauseSharerOptions(**opts))
            _build(single, n, nsolvers)
        # odd steps run without the workers' record export (the default): the hit triples of the other ranks are then
        # not on rank 0 (debugLastHits is a parity hook), the handed-over clauses must still be the single-device ones
        runner = mgpu.PeerRunner(sh, dist, rank, world, payload_cap=8 << 20, slot_hits=1 << 18, records=True)
        res = []
        for step in range(steps):
            if rank == 0:
                _push([sh, single], step, n, nsolvers)
            runner.set_records(step % 2 == 0)
            dist.barrier()
            got = runner.step()
            if rank == 0:
                single.gpuRun(); single.gpuRun()
                want = single.debugLastHits()
                assert got == len(want), (step, got, len(want))
                if step % 2 == 0:
                    mine = sh.debugLastHits()
                    assert np.array_equal(mine, want), step
                a, b = _pop_all(sh, nsolvers), _pop_all(single, nsolvers)
                assert a == b, step
                res.append((int(got), len(a)))
            else:
                # a worker reports only clauses of its own tiles, and some of them
                res.append((int(got), 0))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok", res))
    except BaseException:
        q.put((rank, "error", traceback.format_exc()))


def _run_world(world, n, nsolvers, steps, wait_mode="memop"):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() * 7 + world * 13 + len(wait_mode)) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, nsolvers, steps, wait_mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    try:
        for _ in range(world):
            rank, status, payload = q.get(timeout=240)
            out[rank] = (status, payload)
    finally:
        for p in procs:
            p.join(timeout=20)
            if p.is_alive():
                p.kill()
    for rank in range(world):
        assert rank in out, f"rank {rank} did not report"
        status, payload = out[rank]
        if status == "error" and "DevicesUnavailable" in payload:
            pytest.skip("the device is in exclusive-process mode: a second process cannot share it")
        assert status == "ok", f"rank {rank}:\n{payload}"
    return {r: out[r][1] for r in out}


def test_peer_world1_identical_to_plain_run():
    """world = 1: the window / finalize / import-from-window path alone, in-process"""
    from gpusharesat_b200 import GpuClauseSharer, GpuClauseSharerOptions, mgpu
    n, nsolvers = 15000, 3
    opts = dict(minGpuLatencyMicros=0, verbosity=0)
    sh, single = GpuClauseSharer(GpuClauseSharerOptions(**opts)), GpuClauseSharer(GpuClauseSharerOptions(**opts))
    sh.setShard(0, 1)
    _build(sh, n, nsolvers)
    _build(single, n, nsolvers)
    runner = mgpu.PeerRunner(sh, None, 0, 1, payload_cap=4 << 20, slot_hits=1 << 17)
    big = False
    for step in range(4):
        _push([sh, single], step, n, nsolvers)
        got = runner.step()
        single.gpuRun(); single.gpuRun()
        want = single.debugLastHits()
        assert got == len(want) and len(want) > 0
        big = big or len(want) >= 8192
        assert np.array_equal(sh.debugLastHits(), want)
        assert _pop_all(sh, nsolvers) == _pop_all(single, nsolvers)
        assert runner.device_us() > 0
    assert big  # the device-side sort / resolve of a large union was exercised


@pytest.mark.parametrize("wait_mode", ["memop", "kernel"])
def test_peer_two_processes_identical_to_unsharded(wait_mode):
    """two ranks = two processes on the one device.  n = 24000 makes the first batches overflow the
    initial survivor buffers (second-round protocol) and the unions large enough for the device-side
    sort / resolve on rank 0."""
    res = _run_world(2, 24000, 4, 4, wait_mode)
    root, worker = res[0], res[1]
    assert all(h > 0 and p > 0 for h, p in root)
    # the worker found hits of its own, and fewer than the union
    assert all(0 < w[0] < r[0] for w, r in zip(worker, root))


def test_peer_three_processes_small_db():
    """three ranks, a database so small that some rank owns no tile of some length"""
    res = _run_world(3, 500, 2, 3)
    assert all(h > 0 for h, _ in res[0])
