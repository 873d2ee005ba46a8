"""The reference's only numeric golden value for the hot path: perftest/perfTest.cu:153-207
asserts gpuReports == 143 (debug build, 15 sweeps) and == 19739 (release, 2000 sweeps) for
2 000 000 clauses of length 12-19 over 500 variables, one assignment per sweep."""
import numpy as np
import pytest

from gpusharesat_b200 import GpuClauseSharer, GpuClauseSharerOptions, GlobalStats, mkLit
from oracle_lib import KatAssignments, kat_clauses

pytestmark = pytest.mark.gpu


def _kat(n):
    nvars = 500
    offsets, lits = kat_clauses(2000000, 12, 20, nvars)
    # perfTest.cu:44-54: 10 blocks x 1024 threads (the threads guideline is clamped to 256 here)
    sh = GpuClauseSharer(GpuClauseSharerOptions(gpuBlockCountGuideline=10, gpuThreadsPerBlockGuideline=1024,
                                                initReportCountPerCategory=2000, minGpuLatencyMicros=0))
    sh.setVarCount(nvars)
    sh.setCpuSolverCount(1)
    assert sh.addClausesBulk(offsets, lits) == 0
    stream = KatAssignments(nvars)
    cur = np.full(nvars, 2, dtype=np.uint8)
    popcount = 0
    for _ in range(n):
        vals = stream.next()
        # resetAllVariables (perfTest.cu:79-84): cancelUntil(0) then re-enqueue -> unset + set
        ch = np.nonzero(vals != cur)[0]
        unset = [mkLit(int(v)) for v in ch if vals[v] == 2]
        sets = [mkLit(int(v), bool(vals[v] == 1)) for v in ch if vals[v] != 2]
        sh.unsetSolverValues(0, unset)
        assert sh.trySetSolverValues(0, sets)
        assert sh.trySendAssignment(0) >= 0
        cur = vals
        sh.gpuRun(); sh.gpuRun()  # execute()
        popcount += int(sum(bin(int(m)).count("1") for m in sh.debugLastHits()["mask"]))
        while sh.popReportedClause(0) is not None:
            pass
    return sh.getGlobalStat(GlobalStats.gpuReports), popcount


def test_perftest_kat_debug_count():
    reports, pop = _kat(15)
    assert reports == 143 and pop == 143


def test_perftest_kat_release_count():
    reports, pop = _kat(2000)
    assert reports == 19739 and pop == 19739
