"""reduceDb and the re-sort of streamed clauses on the device (csrc/reduce.cu) against the host path
(host compaction + counting sort + re-upload, GPUSHARE_HOST_REDUCE) fed the same calls: same surviving
clauses, same hits, same hand-over ORDER (the order of the arenas is observable through it), same
statistics -- and both against the oracle's snapshot model for the hits.  Also: arenas that grow in place
(csrc/vmem.cc) keep their contents across many growth steps.
Reference behaviour being matched: Clauses.cu:426-465 (reduceDb), :249-282 (removeClauses), :492-525."""
import os

import numpy as np
import pytest

from gpusharesat_b200 import GlobalStats, GpuClauseSharer, GpuClauseSharerOptions, mkLit
from oracle_lib import SharerModel

pytestmark = pytest.mark.gpu


def make(host_reduce, resort_min=None, **env):
    keys = {"GPUSHARE_HOST_REDUCE": "1" if host_reduce else None,
            "GPUSHARE_RESORT_MIN_CLAUSES": str(resort_min) if resort_min else None}
    keys.update(env)
    old = {k: os.environ.get(k) for k in keys}
    try:
        for k, v in keys.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        return GpuClauseSharer(GpuClauseSharerOptions(minGpuLatencyMicros=0, verbosity=0))
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def pop_all(sh, s):
    out = []
    while (r := sh.popReportedClause(s)) is not None:
        out.append((r[1], tuple(r[0])))
    return out


def add_clauses(shs, rng, nvars, count, lo=1, hi=9):
    for _ in range(count):
        n = int(rng.integers(lo, hi))
        lits = [mkLit(int(rng.integers(0, nvars)), bool(rng.integers(0, 2))) for _ in range(n)]
        assert len({sh.addClause(-1, lits) for sh in shs}) == 1


def push(shs, rng, nvars, nsolvers):
    for s in range(nsolvers):
        for _ in range(int(rng.integers(1, 6))):
            vs = rng.choice(nvars, size=int(rng.integers(nvars // 4, nvars // 2)), replace=False)
            x = rng.random(len(vs))
            unset = [mkLit(int(v)) for v, xx in zip(vs, x) if xx < 0.15]
            sets = [mkLit(int(v), bool(xx < 0.8)) for v, xx in zip(vs, x) if xx >= 0.15]
            for sh in shs:
                sh.unsetSolverValues(s, unset)
            assert len({sh.trySetSolverValues(s, sets) for sh in shs}) == 1
            assert len({sh.trySendAssignment(s) for sh in shs}) == 1


def run_and_compare(a, b, nsolvers, tag):
    for sh in (a, b):
        sh.gpuRun()
        sh.gpuRun()
    ha, hb = a.debugLastHits(), b.debugLastHits()
    assert np.array_equal(ha, hb), tag
    for s in range(nsolvers):
        assert pop_all(a, s) == pop_all(b, s), (tag, s)
    return len(ha)


@pytest.mark.parametrize("max_len", [None, 6])
def test_device_reduce_equals_host_reduce(max_len, tmp_path):
    """max_len = 6: clauses of exactly the maximum length exist and are always removed (Clauses.cu:255)"""
    rng = np.random.default_rng(5 if max_len else 4)
    nvars, nsolvers = 90, 3
    a, b = make(False), make(True)
    for sh in (a, b):
        if max_len:
            sh.setMaxClauseLen(max_len)
        sh.setVarCount(nvars)
        sh.setCpuSolverCount(nsolvers)
    hi = (max_len + 1) if max_len else 9
    total = 0
    for rnd in range(6):
        add_clauses((a, b), rng, nvars, 1500, 1, hi)
        push((a, b), rng, nvars, nsolvers)
        total += run_and_compare(a, b, nsolvers, ("before", rnd))  # hits bump the activities on the device
        before = a.getGlobalStat(GlobalStats.gpuClauses)
        for sh in (a, b):
            sh.reduceDb()
        for st in (GlobalStats.gpuClauses, GlobalStats.gpuClauseLengthSum, GlobalStats.gpuReduceDbs):
            assert a.getGlobalStat(st) == b.getGlobalStat(st), (rnd, st)
        assert a.getGlobalStat(GlobalStats.gpuClauses) < before
        assert a.debugDbOrder()[0] == 0  # every arena is in first-literal order again
        push((a, b), rng, nvars, nsolvers)
        total += run_and_compare(a, b, nsolvers, ("after", rnd))
        # the host mirror was refreshed from the device (asynchronously): the CNF dumps must be the same text
        pa, pb = str(tmp_path / "a.cnf"), str(tmp_path / "b.cnf")
        a.writeClausesInCnf(pa)
        b.writeClausesInCnf(pb)
        assert open(pa).read() == open(pb).read(), rnd
    assert total > 0
    for st in range(9):
        assert a.getGlobalStat(st) == b.getGlobalStat(st), st


def test_streamed_clauses_are_resorted_on_the_device():
    """clauses streamed in behind the sorted part of their arena trigger a device-side re-sort once they are an
    eighth of the database; the hits stay those of the oracle, ids survive the permutation"""
    rng = np.random.default_rng(9)
    nvars, nsolvers = 120, 2
    a, b = make(False, resort_min=200), make(True)
    model = SharerModel(nvars, nsolvers)
    for sh in (a, b):
        sh.setVarCount(nvars)
        sh.setCpuSolverCount(nsolvers)
    total = 0
    for rnd in range(12):
        for _ in range(700):
            n = int(rng.integers(1, 7))
            lits = [mkLit(int(rng.integers(0, nvars)), bool(rng.integers(0, 2))) for _ in range(n)]
            assert len({sh.addClause(-1, lits) for sh in (a, b)}) == 1
            model.addClause(lits)
        push((a, b, model), rng, nvars, nsolvers)
        for sh in (a, b):
            sh.gpuRun()
            sh.gpuRun()
        want = model.run()
        model.run()
        ha, hb = a.debugLastHits(), b.debugLastHits()
        assert np.array_equal(ha, want), rnd
        assert np.array_equal(hb, want), rnd
        total += len(ha)
        for s in range(nsolvers):
            # (the two differ in arena order, hence in hand-over order: same clauses)
            assert sorted(pop_all(a, s)) == sorted(pop_all(b, s)), (rnd, s)
    assert total > 0
    unsorted, resorts = a.debugDbOrder()
    assert resorts >= 3 and unsorted < 8400 // 4
    assert b.debugDbOrder()[1] == 0
    # a reduceDb on top: both arrive at the same database
    for sh in (a, b):
        sh.reduceDb()
    assert a.getGlobalStat(GlobalStats.gpuClauses) == b.getGlobalStat(GlobalStats.gpuClauses)
    push((a, b), rng, nvars, nsolvers)
    assert run_and_compare(a, b, nsolvers, "after reduce") > 0


def test_arenas_grow_in_place_and_keep_their_contents():
    """binary clauses only, 3 M of them (24 MB of ids, 24 MB of literals: past the 8 MB bound where a buffer
    moves behind a virtual address range), added in steps so that the arena grows many times between runs;
    every step's hits must be those of a database loaded in one piece"""
    rng = np.random.default_rng(2)
    nvars, nsolvers, steps, per_step = 4000, 2, 6, 500_000
    a = make(False)
    b = make(False, GPUSHARE_NO_VMM="1")
    for sh in (a, b):
        sh.setVarCount(nvars)
        sh.setCpuSolverCount(nsolvers)
    for step in range(steps):
        v = rng.integers(0, nvars, size=(per_step, 2))
        sg = rng.random(size=(per_step, 2)) < 0.1  # mostly positive literals, mostly true below: ~3 % of the clauses hit
        lits = (2 * v + sg).astype(np.int32).reshape(-1)
        off = np.arange(per_step + 1, dtype=np.int64) * 2
        for sh in (a, b):
            sh.addClausesBulk(off, lits)
        for s in range(nsolvers):
            vs = rng.choice(nvars, size=nvars - 40, replace=False)
            sets = (2 * vs + (rng.random(len(vs)) < 0.1)).astype(np.int32)
            for sh in (a, b):
                sh.unsetSolverValues(s, [mkLit(int(x)) for x in range(0, nvars, 97)])
                assert sh.trySetSolverValues(s, sets)
                assert sh.trySendAssignment(s) >= 0
        for sh in (a, b):
            sh.gpuRun()
            sh.gpuRun()
        ha, hb = a.debugLastHits(), b.debugLastHits()
        assert len(ha) > 0 and np.array_equal(ha, hb), step
        for s in range(nsolvers):
            assert sorted(pop_all(a, s)) == sorted(pop_all(b, s))
    assert a.getGlobalStat(GlobalStats.gpuClauses) == steps * per_step
