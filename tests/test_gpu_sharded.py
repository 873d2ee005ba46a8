"""Clause sharding (multi-GPU path) on ONE device: two sharers hold the two halves of the tile
space and run the same broadcast payload; the union of their hits must equal the unsharded run and
the CPU oracle, and rank 0's hand-over must behave like the single-GPU one."""
import ctypes as C

import numpy as np
import pytest

from gpusharesat_b200 import GpuClauseSharer, GpuClauseSharerOptions, mkLit
from oracle_lib import SharerModel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("dense", [False, True])
def test_sharded_union_equals_unsharded_and_oracle(world, dense):
    rng = np.random.default_rng(40 + world)
    nvars, nsolvers = 60, 5
    opts = dict(gpuBlockCountGuideline=3, gpuThreadsPerBlockGuideline=64, minGpuLatencyMicros=0, initReportCountPerCategory=3)
    ranks = [GpuClauseSharer(GpuClauseSharerOptions(**opts)) for _ in range(world)]
    single = GpuClauseSharer(GpuClauseSharerOptions(**opts))
    model = SharerModel(nvars, nsolvers)
    for r, sh in enumerate(ranks):
        sh.setShard(r, world)
    for sh in ranks + [single]:
        sh.setVarCount(nvars)
        sh.setCpuSolverCount(nsolvers)
        sh.debugSetDense(dense)
    total = 0
    for rnd in range(6):
        for _ in range(int(rng.integers(100, 700))):  # several tiles per length so every rank owns some
            n = int(rng.integers(1, 6))
            lits = [mkLit(int(rng.integers(0, nvars)), bool(rng.integers(0, 2))) for _ in range(n)]
            ids = {sh.addClause(-1, lits) for sh in ranks + [single]}
            assert ids == {model.addClause(lits)}
        front = ranks[0]
        for s in range(nsolvers):
            for _ in range(int(rng.integers(0, 12))):
                vs = rng.choice(nvars, size=int(rng.integers(0, 30)), replace=False)
                x = rng.random(len(vs))
                unset = [mkLit(int(v)) for v, xx in zip(vs, x) if xx < 0.2]
                sets = [mkLit(int(v), bool(xx < 0.85)) for v, xx in zip(vs, x) if xx >= 0.2]
                for sh in (front, single, model):
                    sh.unsetSolverValues(s, unset)
                    assert sh.trySetSolverValues(s, sets)
                    assert sh.trySendAssignment(s) >= 0
        # rank 0 collects; every rank runs the same payload (host pointers here, NCCL-broadcast
        # device buffers in bench.py); hits are gathered and handed over on rank 0
        rebuild, pptr, pbytes, uptr, nupd = front.mgpuCollect()
        parts = []
        for sh in ranks:
            sh.mgpuRun(pptr, pbytes, uptr, nupd, rebuild)
        for sh in ranks:
            parts.append(sh.mgpuWait())
        for r, p in enumerate(parts):  # a rank only reports clauses of its own tiles
            assert np.all((p["idx"] // 128) % world == r)
        union = np.concatenate(parts)
        front.mgpuImport(union)
        got = front.debugLastHits()
        single.gpuRun(); single.gpuRun()
        want = model.run(); model.run()
        assert np.array_equal(got, want), rnd
        assert np.array_equal(single.debugLastHits(), want)
        total += len(want)
        for s in range(nsolvers):
            a, b = [], []
            while (x := front.popReportedClause(s)) is not None:
                a.append(x)
            while (x := single.popReportedClause(s)) is not None:
                b.append(x)
            assert a == b
    assert total > 0


def test_packed_payload_fast_path_single_device():
    """gss_mgpu_collect_to / gss_mgpu_run_payload / gss_mgpu_hits_to_device: the payload is copied
    device-to-device between two sharers instead of being broadcast."""
    import torch
    rng = np.random.default_rng(99)
    nvars, nsolvers, world = 80, 3, 2
    opts = dict(minGpuLatencyMicros=0)
    ranks = [GpuClauseSharer(GpuClauseSharerOptions(**opts)) for _ in range(world)]
    model = SharerModel(nvars, nsolvers)
    for r, sh in enumerate(ranks):
        sh.setShard(r, world)
        sh.setVarCount(nvars)
        sh.setCpuSolverCount(nsolvers)
    dev = torch.device("cuda", 0)
    bufs = [torch.zeros(1 << 20, dtype=torch.uint8, device=dev) for _ in range(world)]
    hitbuf = torch.zeros(1 << 20, dtype=torch.uint8, device=dev)
    from gpusharesat_b200.api import RAW_HIT_DTYPE
    # no clause yet: the header says "nothing to run"
    n = ranks[0].mgpuCollectTo(bufs[0].data_ptr(), bufs[0].numel())
    torch.cuda.synchronize()
    assert n == 64 and ranks[0].mgpuRunPayload(bufs[0].data_ptr(), bufs[0].numel()) == -1
    for rnd in range(4):
        for _ in range(600):
            k = int(rng.integers(1, 5))
            lits = [mkLit(int(rng.integers(0, nvars)), bool(rng.integers(0, 2))) for _ in range(k)]
            assert {sh.addClause(-1, lits) for sh in ranks} == {model.addClause(lits)}
        for s in range(nsolvers):
            for _ in range(5):
                vs = rng.choice(nvars, size=40, replace=False)
                sets = [mkLit(int(v), bool(rng.random() < 0.85)) for v in vs]
                for sh in (ranks[0], model):
                    assert sh.trySetSolverValues(s, sets) and sh.trySendAssignment(s) >= 0
        total = ranks[0].mgpuCollectTo(bufs[0].data_ptr(), bufs[0].numel())
        torch.cuda.synchronize()
        bufs[1][:total].copy_(bufs[0][:total])  # stands in for the NCCL broadcast
        torch.cuda.synchronize()
        parts = []
        for r, sh in enumerate(ranks):
            assert sh.mgpuRunPayload(bufs[r].data_ptr(), bufs[r].numel()) == (1 if rnd == 0 else 0)
        for sh in ranks:
            cnt = sh.mgpuWaitCount()
            got = sh.mgpuHitsToDevice(hitbuf.data_ptr(), cnt)
            torch.cuda.synchronize()
            assert got == cnt
            parts.append(hitbuf[: cnt * 16].cpu().numpy().view(RAW_HIT_DTYPE).copy())
        ranks[0].mgpuImport(np.concatenate(parts))
        want = model.run(); model.run()
        assert np.array_equal(ranks[0].debugLastHits(), want) and len(want) > 0
