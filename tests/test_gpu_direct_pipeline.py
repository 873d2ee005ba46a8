"""The direct run pipeline (csrc/pipeline.cu: deltas read by the GPU from the solver threads' own
page-locked buffers, per-solver record lists sorted and emitted by k_emit into page-locked result
buffers that the solvers' batches view in place) against the staged pipeline it replaces
(GPUSHARE_LEGACY_PIPELINE=1: staging copy + H2D, one global hit buffer, host / CUB sort, literal
copies): identical hit triples, identical pop sequences, identical statistics -- on random traffic,
on record lists too long for k_emit's shared-memory sort, with 70 solvers, and with a solver that
never pops (the result-buffer safety valve)."""
import os

import numpy as np
import pytest

from gpusharesat_b200 import GpuClauseSharer, GpuClauseSharerOptions, mkLit
from oracle_lib import SharerModel

pytestmark = pytest.mark.gpu


def make(legacy, **opts):
    if legacy:
        os.environ["GPUSHARE_LEGACY_PIPELINE"] = "1"
    else:
        os.environ.pop("GPUSHARE_LEGACY_PIPELINE", None)
    try:
        return GpuClauseSharer(GpuClauseSharerOptions(minGpuLatencyMicros=0, **opts))
    finally:
        os.environ.pop("GPUSHARE_LEGACY_PIPELINE", None)


def pop_all(sh, s):
    out = []
    while True:
        r = sh.popReportedClause(s)
        if r is None:
            return out
        out.append((r[1], tuple(r[0])))


def drive_pair(nvars, nsolvers, nrounds, seed, max_len=8, clauses_per_round=60, sends=12, **opts):
    rng = np.random.default_rng(seed)
    a, b = make(False, **opts), make(True, **opts)
    model = SharerModel(nvars, nsolvers)
    for sh in (a, b):
        sh.setVarCount(nvars)
        sh.setCpuSolverCount(nsolvers)
    total = 0
    for r in range(nrounds):
        for _ in range(clauses_per_round):
            n = int(rng.integers(1, max_len + 1))
            lits = [mkLit(int(rng.integers(0, nvars)), bool(rng.integers(0, 2))) for _ in range(n)]
            s = int(rng.integers(-1, nsolvers))
            ids = {sh.addClause(s, lits) for sh in (a, b)}
            model.addClause(lits)
            assert len(ids) == 1
        for s in range(nsolvers):
            for _ in range(int(rng.integers(0, sends))):
                vs = rng.choice(nvars, size=int(rng.integers(1, nvars // 2)), replace=False)
                x = rng.random(len(vs))
                unset = [mkLit(int(v)) for v, xx in zip(vs, x) if xx < 0.2]
                sets = [mkLit(int(v), bool(xx < 0.75)) for v, xx in zip(vs, x) if xx >= 0.2]
                for sh in (a, b, model):
                    sh.unsetSolverValues(s, unset)
                oks = {sh.trySetSolverValues(s, sets) for sh in (a, b, model)}
                assert len(oks) == 1
                assert len({sh.trySendAssignment(s) for sh in (a, b, model)}) == 1
        for sh in (a, b):
            sh.gpuRun()
            sh.gpuRun()
        expect = model.run()
        model.run()
        ha, hb = a.debugLastHits(), b.debugLastHits()
        assert np.array_equal(ha, hb), r
        if expect is not None:
            assert np.array_equal(ha, expect), r
        total += len(ha)
        for s in range(nsolvers):
            if rng.random() < 0.8:  # sometimes a solver leaves its batches for later
                assert pop_all(a, s) == pop_all(b, s), (r, s)
            assert a.getLastAssigAllReported(s) == b.getLastAssigAllReported(s)
    for s in range(nsolvers):
        assert pop_all(a, s) == pop_all(b, s)
        for st in range(6):
            assert a.getOneSolverStat(s, st) == b.getOneSolverStat(s, st), (s, st)
    for st in (0, 1, 2, 3, 4, 5, 6, 8):  # everything but the time gauges
        assert a.getGlobalStat(st) == b.getGlobalStat(st), st
    return total


@pytest.mark.parametrize("nsolvers", [1, 3, 32])
def test_direct_equals_staged_pipeline(nsolvers):
    assert drive_pair(60, nsolvers, 10, seed=300 + nsolvers) > 0


def test_direct_equals_staged_with_70_solvers_and_tiny_first_buffers():
    assert drive_pair(40, 70, 5, seed=17, initReportCountPerCategory=1, gpuBlockCountGuideline=2, sends=6) > 0


def test_record_lists_longer_than_the_shared_memory_sort():
    """30 000 clauses of 1-3 literals all false for two of three solvers: ~30 000 records per solver,
    past k_emit's 8192-record shared-memory sort (the same network then runs in global memory)"""
    rng = np.random.default_rng(5)
    nvars, nsolvers, ncl = 500, 3, 30000
    a, b = make(False), make(True)
    clauses = []
    for _ in range(ncl):
        n = int(rng.integers(1, 4))
        clauses.append([mkLit(int(v)) for v in rng.choice(nvars, size=n, replace=False)])
    for sh in (a, b):
        sh.setVarCount(nvars)
        sh.setCpuSolverCount(nsolvers)
        for c in clauses:
            sh.addClause(-1, c)
        for s in (0, 2):
            assert sh.trySetSolverValues(s, [mkLit(v, True) for v in range(nvars)])  # everything false
            assert sh.trySendAssignment(s) == 0
        sh.gpuRun()
        sh.gpuRun()
    ha, hb = a.debugLastHits(), b.debugLastHits()
    assert len(ha) == 2 * ncl and np.array_equal(ha, hb)
    for s in range(nsolvers):
        pa = pop_all(a, s)
        assert pa == pop_all(b, s)
        assert len(pa) == (0 if s == 1 else ncl)
        # hand-over order: by length, then by position in the length's arena
        assert [len(l) for _, l in pa] == sorted(len(l) for _, l in pa)


def test_solver_that_never_pops_does_not_pin_result_buffers_forever():
    """more than 64 runs whose batches solver 1 never takes: past that bound the slices are copied
    out of the result buffers (safety valve); everything is still delivered, in order"""
    nvars, runs = 20, 90
    sh = make(False)
    sh.setVarCount(nvars)
    sh.setCpuSolverCount(2)
    expect = []
    for r in range(runs):
        cid = sh.addClause(-1, [mkLit(r % nvars), mkLit((r + 1) % nvars)])
        for s in (0, 1):
            sh.unsetSolverValues(s, [mkLit(v) for v in range(nvars)])
            assert sh.trySetSolverValues(s, [mkLit(r % nvars, True), mkLit((r + 1) % nvars, True)])
            assert sh.trySendAssignment(s) == r
        sh.gpuRun()
        sh.gpuRun()
        got0 = pop_all(sh, 0)
        assert (cid, (mkLit(r % nvars), mkLit((r + 1) % nvars))) in got0
        expect.append(cid)
    got1 = [cid for cid, _ in pop_all(sh, 1)]
    assert set(expect) <= set(got1)
