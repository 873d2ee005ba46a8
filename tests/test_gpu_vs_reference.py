"""Bit-exact comparison with the reference's OWN GPU checker: the unmodified gpuShareLib compiled
for sm_100a (oracle/_ref/libgpushare_ref.so, built by oracle/Makefile from /root/reference where
it lies) is driven with the same calls as libgpushare_b200.so; sorted (clause, solver, mask)
triples, per-solver pop sequences and stats must be identical.  Also the BASELINE.json config-2
shape (1 M clauses x 32 assignments over 50 k variables) against reference and CPU oracle."""
import numpy as np
import pytest

import ref_lib
import synth
from gpusharesat_b200 import GpuClauseSharer, GpuClauseSharerOptions, GlobalStats, OneSolverStats, mkLit
from oracle_lib import check_db

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref not built")]


def both(nvars, nsolvers, blocks=3, threads=32, report=5000):
    mine = GpuClauseSharer(GpuClauseSharerOptions(gpuBlockCountGuideline=blocks, gpuThreadsPerBlockGuideline=threads,
                                                  minGpuLatencyMicros=0, initReportCountPerCategory=report))
    ref = ref_lib.RefSharer(blocks=blocks, threads=threads, report=report)
    for s in (mine, ref):
        s.setVarCount(nvars)
        s.setCpuSolverCount(nsolvers)
    return mine, ref


def pop_all(sh, s):
    out = []
    while True:
        r = sh.popReportedClause(s)
        if r is None:
            return out
        out.append((r[1], tuple(r[0])))


@pytest.mark.parametrize("nsolvers", [1, 2, 3, 4, 8, 16, 32, 7])
def test_random_traffic_identical_to_reference(nsolvers):
    # nsolvers == 7 exposes a reference defect: DAssigAggregates::getEndBitPos (Assigs.cuh:68-74)
    # omits lowBitsStart for the solvers that own the smaller bit groups, so when 32 % nsolvers != 0
    # dCheckOneClauseAllSolvers (GpuRunner.cu:116-131) stops after the first such solver and the
    # later solvers' hits are lost.  There the reference must report a SUBSET of ours (ours is
    # checked against the CPU oracle in test_gpu_random_parity.py).
    subset_only = 32 % nsolvers != 0 and nsolvers != 3
    rng = np.random.default_rng(1000 + nsolvers)
    nvars = 50
    mine, ref = both(nvars, nsolvers)
    total = 0
    for r in range(10):
        for _ in range(int(rng.integers(0, 60))):
            n = int(rng.integers(1, 10))
            lits = [mkLit(int(rng.integers(0, nvars)), bool(rng.integers(0, 2))) for _ in range(n)]
            src = int(rng.integers(-1, nsolvers))
            assert mine.addClause(src, lits) == ref.addClause(src, lits)
        for s in range(nsolvers):
            for _ in range(int(rng.integers(0, 40))):
                vs = rng.choice(nvars, size=int(rng.integers(0, 20)), replace=False)
                x = rng.random(len(vs))
                unset = [mkLit(int(v)) for v, xx in zip(vs, x) if xx < 0.25]
                sets = [mkLit(int(v), bool(xx < 0.8)) for v, xx in zip(vs, x) if xx >= 0.25]
                mine.unsetSolverValues(s, unset); ref.unsetSolverValues(s, unset)
                assert mine.trySetSolverValues(s, sets) == ref.trySetSolverValues(s, sets)
                assert mine.trySendAssignment(s) == ref.trySendAssignment(s)
            assert np.array_equal(mine.getCurrentAssignment(s, nvars), ref.getCurrentAssignment(s, nvars))
        for sh in (mine, ref):
            sh.gpuRun(); sh.gpuRun()
        a, b = mine.debugLastHits(), ref.debugLastHits()
        if subset_only:
            assert set(map(tuple, b.tolist())) <= set(map(tuple, a.tolist()))
            total += len(a)
            continue
        assert np.array_equal(a, b), (r, len(a), len(b))
        total += len(a)
        for s in range(nsolvers):
            if rng.random() < 0.7:
                pa, pb = pop_all(mine, s), pop_all(ref, s)
                nhits = int(np.sum(a["solver_id"] == s))
                # A re-reported clause ends its batch in the reference (Reported.cu:113-129), so what
                # is handed over then depends on the order the GPU happened to emit the hits in.
                # Without such a duplicate the hand-over is order independent and must be identical.
                if len(pa) == nhits and len(pb) == nhits:
                    assert sorted(pa) == sorted(pb)
                assert mine.getLastAssigAllReported(s) == ref.getLastAssigAllReported(s)
    for st in (GlobalStats.gpuClauses, GlobalStats.gpuClauseLengthSum, GlobalStats.gpuClausesAdded,
               GlobalStats.gpuRuns, GlobalStats.gpuReports, GlobalStats.clauseTestsOnGroups,
               GlobalStats.totalAssigClauseTested):
        if subset_only and st == GlobalStats.gpuReports:
            continue
        assert mine.getGlobalStat(st) == ref.getGlobalStat(st), st
    for s in range(nsolvers):
        for st in (OneSolverStats.varUpdatesSentToGpu, OneSolverStats.assigsSentToGpu, OneSolverStats.failuresToFindAssig):
            assert mine.getOneSolverStat(s, st) == ref.getOneSolverStat(s, st), (s, st)
    assert total > 0


def _config2(nclauses, nvars, nslots, churn):
    sig = synth.sigma(nvars, 11)
    offsets, lits = synth.clauses(nclauses, nvars, 30, sig, 0.98, 12)
    stream = synth.Stream(nvars, sig, 0.01, churn, 13)
    deltas, snaps = [], []
    for _ in range(nslots):
        s, u = stream.next()
        deltas.append((s.copy(), u.copy()))
        snaps.append(stream.values())
    return offsets, lits, deltas, snaps


def _drive(sh, offsets, lits, deltas):
    sh.addClausesBulk(offsets, lits)
    for sets, unsets in deltas:
        sh.unsetSolverValues(0, unsets)
        assert sh.trySetSolverValues(0, sets)
        assert sh.trySendAssignment(0) >= 0
    sh.gpuRun(); sh.gpuRun()
    return sh.debugLastHits()


def test_config2_shape_bit_exact_vs_reference_and_oracle():
    # BASELINE.json configs[1]: 1M clauses (len 2-30, Luby-like mix) x 32 assignments over 50k vars
    nclauses, nvars = 1000000, 50000
    offsets, lits, deltas, snaps = _config2(nclauses, nvars, 32, 0.01)
    mine = GpuClauseSharer(GpuClauseSharerOptions(minGpuLatencyMicros=0))
    mine.setVarCount(nvars); mine.setCpuSolverCount(1)
    got = _drive(mine, offsets, lits, deltas)
    d, t, start = synth.pack_slots(snaps)
    want = check_db(offsets, lits, d[None, :], t[None, :], np.array([start], dtype=np.uint32), use_filter=1, nthreads=8)
    assert len(want) > 100
    assert np.array_equal(got, want)
    # dense mode (bench-only kernel): identical hit set, both as the run's own mode and when the
    # check is re-launched on the tables of a started run (gss_debug_time_check)
    dense = GpuClauseSharer(GpuClauseSharerOptions(minGpuLatencyMicros=0))
    dense.setVarCount(nvars); dense.setCpuSolverCount(1)
    dense.debugSetDense(True)
    assert np.array_equal(_drive(dense, offsets, lits, deltas), want)
    assert dense.trySendAssignment(0) >= 0  # one more slot holding the last assignment
    dense.debugSetDense(False)
    dense.gpuRun()
    assert dense.debugTimeCheck(2, dense=True) > 0
    dense.gpuRun()
    again = dense.debugLastHits()
    d1, t1, _ = synth.pack_slots([snaps[-1]])
    want1 = check_db(offsets, lits, d1[None, :], t1[None, :], np.array([1], dtype=np.uint32), use_filter=0, nthreads=8)
    assert np.array_equal(again, want1) and len(want1) > 0
    # the reference's own GPU checker (hit buffer large enough that it cannot drop hits)
    ref = ref_lib.RefSharer(report=20000)
    ref.setVarCount(nvars); ref.setCpuSolverCount(1)
    assert np.array_equal(_drive(ref, offsets, lits, deltas), want)


def test_reduce_db_identical_to_reference():
    # ClauseActivityLbdTest.cu / GpuSolverTest.cu:889-935: activity-only halving, lengths 1-2 kept
    rng = np.random.default_rng(77)
    nvars = 60
    mine, ref = both(nvars, 2, report=20000)
    for rnd in range(4):
        for _ in range(600):
            n = int(rng.integers(1, 8))
            lits = [mkLit(int(rng.integers(0, nvars)), bool(rng.integers(0, 2))) for _ in range(n)]
            assert mine.addClause(-1, lits) == ref.addClause(-1, lits)
        for s in range(2):
            for _ in range(6):
                vs = rng.choice(nvars, size=30, replace=False)
                sets = [mkLit(int(v), bool(rng.random() < 0.8)) for v in vs]
                assert mine.trySetSolverValues(s, sets) == ref.trySetSolverValues(s, sets)
                assert mine.trySendAssignment(s) == ref.trySendAssignment(s)
        for sh in (mine, ref):
            sh.gpuRun(); sh.gpuRun()
        assert np.array_equal(mine.debugLastHits(), ref.debugLastHits())
        for sh in (mine, ref):
            sh.reduceDb()
        for st in (GlobalStats.gpuClauses, GlobalStats.gpuClauseLengthSum, GlobalStats.gpuReduceDbs):
            assert mine.getGlobalStat(st) == ref.getGlobalStat(st), (rnd, st)
        for s in range(2):
            pop_all(mine, s), pop_all(ref, s)
    assert mine.getGlobalStat(GlobalStats.gpuClauses) < 2400
