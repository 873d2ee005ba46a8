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
