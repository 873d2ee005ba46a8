"""The reference's own integration fixtures (glucose-syrup/test/GpuSolverTest.cu), re-expressed
through the C ABI (ctypes mirror of GpuClauseSharer).  Boost.Test is not available and the
reference's fixtures poke protected members; here everything goes through the public API."""
import numpy as np
import pytest

from gpusharesat_b200 import GpuClauseSharer, GpuClauseSharerOptions, GlobalStats, OneSolverStats, mkLit

pytestmark = pytest.mark.gpu

TRUE, FALSE, UNDEF = 0, 1, 2


def default_opts(**kw):
    # reference setDefaultOptions, testUtils/TestHelper.cu:31-35: a tiny grid so that warps loop
    o = dict(gpuBlockCountGuideline=3, gpuThreadsPerBlockGuideline=32, minGpuLatencyMicros=50)
    o.update(kw)
    return GpuClauseSharerOptions(**o)


def make(nvars, nsolvers, **kw):
    sh = GpuClauseSharer(default_opts(**kw))
    sh.setVarCount(nvars)
    sh.setCpuSolverCount(nsolvers)
    return sh


def execute(sh):
    # TestHelper.cu:55-59: results of a run surface one gpuRun() later
    sh.gpuRun()
    sh.gpuRun()


def set_vals(sh, s, assign):
    """assign: {var: TRUE/FALSE/UNDEF}"""
    unset = [mkLit(v) for v, x in assign.items() if x == UNDEF]
    sets = [mkLit(v, x == FALSE) for v, x in assign.items() if x != UNDEF]
    if unset:
        sh.unsetSolverValues(s, unset)
    if sets:
        assert sh.trySetSolverValues(s, sets)


def popped(sh, s):
    out = []
    while True:
        r = sh.popReportedClause(s)
        if r is None:
            return out
        out.append(r)


def test_clauses_assigs_reported():
    # GpuSolverTest.cu:343-391 testClausesAssigsReported
    sh = make(3, 3)
    for v in range(3):
        sh.addClause(-1, [mkLit(v)])
    set_vals(sh, 0, {0: FALSE, 1: TRUE, 2: TRUE}); assert sh.trySendAssignment(0) == 0
    set_vals(sh, 0, {0: TRUE, 1: FALSE, 2: FALSE}); assert sh.trySendAssignment(0) == 1
    set_vals(sh, 1, {0: TRUE, 1: FALSE, 2: TRUE}); assert sh.trySendAssignment(1) == 0
    execute(sh)
    assert len(popped(sh, 0)) == 3
    assert len(popped(sh, 1)) == 1
    assert len(popped(sh, 2)) == 0
    set_vals(sh, 0, {1: TRUE}); sh.trySendAssignment(0)
    execute(sh)
    assert len(popped(sh, 0)) == 1
    assert len(popped(sh, 1)) == 0


def test_find_clauses_multi_thread():
    # GpuSolverTest.cu:394-454 testFindClausesMultiThread
    sh = make(3, 1, gpuBlockCountGuideline=1)
    set_vals(sh, 0, {0: FALSE, 1: TRUE, 2: UNDEF}); sh.trySendAssignment(0)
    sh.addClause(-1, [mkLit(0), mkLit(1)])
    sh.gpuRun()  # copy the clauses several times (the reference calls copyToDeviceAsync between adds)
    sh.addClause(-1, [mkLit(0), mkLit(1, True)])
    sh.addClause(-1, [mkLit(1, True), mkLit(2)])
    execute(sh)
    # the first run saw only clause 0 (satisfied); the slot is tested once, so run again
    got = popped(sh, 0)
    assert got == []
    set_vals(sh, 0, {}); sh.trySendAssignment(0)
    execute(sh)
    got = sorted(popped(sh, 0), key=lambda r: r[1])
    assert [g[0] for g in got] == [[mkLit(0), mkLit(1, True)], [mkLit(1, True), mkLit(2)]]
    assert [g[1] for g in got] == [1, 2]
    assert sh.popReportedClause(0) is None


def test_find_clauses_all_added_first():
    # same fixture with all three clauses present before the run: exactly {0,-1} and {-1,2} fire
    sh = make(3, 1, gpuBlockCountGuideline=1)
    sh.addClause(-1, [mkLit(0), mkLit(1)])
    sh.addClause(-1, [mkLit(0), mkLit(1, True)])
    sh.addClause(-1, [mkLit(1, True), mkLit(2)])
    set_vals(sh, 0, {0: FALSE, 1: TRUE, 2: UNDEF}); sh.trySendAssignment(0)
    execute(sh)
    hits = sh.debugLastHits()
    assert hits["clause_id"].tolist() == [1, 2]
    assert hits["mask"].tolist() == [1, 1]
    got = sorted(popped(sh, 0), key=lambda r: r[1])
    assert [g[0] for g in got] == [[mkLit(0), mkLit(1, True)], [mkLit(1, True), mkLit(2)]]


def test_many_clauses_reported_all_at_once():
    # GpuSolverTest.cu:696-724: 4000 unit clauses, every third variable already true
    n = 4000
    sh = make(n, 1, gpuBlockCountGuideline=2, initReportCountPerCategory=5000)
    assign = {}
    for i in range(n):
        if i % 3 == 0:
            assign[i] = TRUE
        sh.addClause(-1, [mkLit(i)])
    set_vals(sh, 0, assign); sh.trySendAssignment(0)
    execute(sh)
    expected = n - (n + 2) // 3
    assert len(sh.debugLastHits()) == expected
    assert len(popped(sh, 0)) == expected
    assert sh.getOneSolverStat(0, OneSolverStats.reportedClauses) == expected
    assert sh.getOneSolverStat(0, OneSolverStats.reportedClausesUnit) == expected
    assert sh.getGlobalStat(GlobalStats.gpuReports) == expected


def test_hit_buffer_overflow_is_rerun_not_dropped():
    # same fixture with a 1-record hit buffer: the reference would drop hits (Reporter.cuh:46-48)
    n = 4000
    sh = make(n, 1, gpuBlockCountGuideline=1, initReportCountPerCategory=1)
    for i in range(n):
        sh.addClause(-1, [mkLit(i)])
    set_vals(sh, 0, {i: FALSE for i in range(0, n, 2)}); sh.trySendAssignment(0)
    execute(sh)
    assert len(sh.debugLastHits()) == n  # false -> conflict, undef -> unit: every clause fires
    assert len(popped(sh, 0)) == n


def test_solver_passes_many_assignments():
    # GpuSolverTest.cu:949-981: 32 slots of one solver, 32 binary clauses -> 32 reports
    sh = make(64, 3, initReportCountPerCategory=100)
    prev = None
    for i in range(32):
        sh.addClause(-1, [mkLit(2 * i, True), mkLit(2 * i + 1)])
        if prev is not None:
            sh.unsetSolverValues(0, [mkLit(prev)])
        assert sh.trySetSolverValues(0, [mkLit(2 * i)])
        assert sh.trySendAssignment(0) == i
        prev = 2 * i
    assert sh.trySendAssignment(0) == -1  # all 32 slots frozen
    assert sh.getOneSolverStat(0, OneSolverStats.failuresToFindAssig) == 1
    execute(sh)
    hits = sh.debugLastHits()
    assert hits["clause_id"].tolist() == list(range(32))
    assert hits["mask"].tolist() == [1 << i for i in range(32)]
    assert len(popped(sh, 0)) == 32
    assert sh.getOneSolverStat(0, OneSolverStats.reportedClausesBinary) == 32


def test_one_assignment_then_two():
    # GpuSolverTest.cu:483-508: all aggregate bits must follow the collapse, not only the used ones
    sh = make(4, 1)
    sh.addClause(-1, [mkLit(0, True), mkLit(1, True), mkLit(2)])
    assert sh.trySetSolverValues(0, [mkLit(0)]); sh.trySendAssignment(0)
    execute(sh)
    assert popped(sh, 0) == []
    assert sh.trySetSolverValues(0, [mkLit(3)]); sh.trySendAssignment(0)
    assert sh.trySetSolverValues(0, [mkLit(1)]); sh.trySendAssignment(0)
    execute(sh)
    assert len(popped(sh, 0)) == 1


def test_doesnt_import_same_clause_twice():
    # GpuSolverTest.cu:511-536: two assignments of one run hit the same clause -> one report
    sh = make(3, 1)
    assert sh.trySetSolverValues(0, [mkLit(0)]); sh.trySendAssignment(0)
    sh.unsetSolverValues(0, [mkLit(0)])
    assert sh.trySetSolverValues(0, [mkLit(1, True)]); sh.trySendAssignment(0)
    sh.addClause(-1, [mkLit(0, True), mkLit(1)])
    execute(sh)
    assert sh.debugLastHits()["mask"].tolist() == [3]
    assert len(popped(sh, 0)) == 1
    assert sh.getOneSolverStat(0, OneSolverStats.reportedClauses) == 1


def test_doesnt_import_same_clause_twice_on_successive_runs():
    # GpuSolverTest.cu:540-563
    sh = make(3, 1)
    sh.addClause(-1, [mkLit(0, True), mkLit(1)])
    assert sh.trySetSolverValues(0, [mkLit(0)]); sh.trySendAssignment(0)
    sh.gpuRun()
    sh.unsetSolverValues(0, [mkLit(0)])
    assert sh.trySetSolverValues(0, [mkLit(1, True)]); sh.trySendAssignment(0)
    sh.gpuRun()
    n = len(popped(sh, 0))
    sh.gpuRun()
    n += len(popped(sh, 0))
    assert n == 1
    assert sh.getOneSolverStat(0, OneSolverStats.reportedClauses) == 1


def test_can_reimport_after_all_assignments_reported():
    # Reported.cu:130-146 / GpuSolverTest.cu:566-603: once every assignment that could not know a
    # reported clause has been fully reported, the clause may be reported again
    sh = make(5, 1)
    sh.addClause(-1, [mkLit(0, True), mkLit(1, True), mkLit(4)])
    assert sh.trySetSolverValues(0, [mkLit(0), mkLit(1)]); sh.trySendAssignment(0)
    execute(sh)
    assert len(popped(sh, 0)) == 1
    # the solver "deleted" the clause; the same trail is sent again
    sh.trySendAssignment(0)
    execute(sh)
    assert len(popped(sh, 0)) == 1
    assert sh.getOneSolverStat(0, OneSolverStats.reportedClauses) == 2


def test_exporter_does_not_get_its_own_clause_back():
    # GpuSolverTest.cu:999-1033 testSendClauseToGpu / Reported.cu:97-103
    sh = make(3, 2)
    for s in (0, 1):
        assert sh.trySetSolverValues(s, [mkLit(0), mkLit(1)]); sh.trySendAssignment(s)
    cid = sh.addClause(0, [mkLit(0, True), mkLit(1, True)])
    assert cid == 0
    execute(sh)
    assert sh.debugLastHits()["solver_id"].tolist() == [0, 1]  # the kernel reports both
    assert popped(sh, 0) == []                                  # the exporter is not told
    assert len(popped(sh, 1)) == 1
    assert sh.getGlobalStat(GlobalStats.gpuClauses) == 1
    assert sh.getGlobalStat(GlobalStats.gpuClauseLengthSum) == 2


def test_two_solvers_import_binary():
    # GpuSolverTest.cu:605-633
    sh = make(3, 2)
    assert sh.trySetSolverValues(0, [mkLit(0)]); sh.trySendAssignment(0)
    assert sh.trySetSolverValues(1, [mkLit(0, True)]); sh.trySendAssignment(1)
    sh.addClause(-1, [mkLit(0, True), mkLit(1)])
    sh.addClause(-1, [mkLit(0), mkLit(1, True)])
    execute(sh)
    assert [r[1] for r in popped(sh, 0)] == [0]
    assert [r[1] for r in popped(sh, 1)] == [1]


def test_solver_unsets():
    # GpuSolverTest.cu:636-658
    sh = make(2, 1)
    assert sh.trySetSolverValues(0, [mkLit(1)]); sh.trySendAssignment(0)
    execute(sh)
    assert popped(sh, 0) == []
    sh.unsetSolverValues(0, [mkLit(1)])
    assert sh.trySetSolverValues(0, [mkLit(0)]); sh.trySendAssignment(0)
    sh.addClause(-1, [mkLit(0, True), mkLit(1)])
    execute(sh)
    assert len(popped(sh, 0)) == 1


def test_oversize_clause_rejected_and_limit_can_be_lifted():
    # Clauses.cu:350: > MAX_CL_SIZE (100) -> -1; config 5 needs 200
    sh = make(300, 1)
    assert sh.addClause(-1, [mkLit(v) for v in range(101)]) == -1
    assert sh.addClause(-1, [mkLit(v) for v in range(100)]) == 0
    sh2 = make(300, 1)
    sh2.setMaxClauseLen(200)
    assert sh2.addClause(-1, [mkLit(v) for v in range(200)]) == 0
    assert sh2.addClause(-1, [mkLit(v) for v in range(201)]) == -1
    assert sh2.trySetSolverValues(0, [mkLit(v, True) for v in range(199)]); sh2.trySendAssignment(0)
    execute(sh2)
    assert sh2.debugLastHits()["mask"].tolist() == [1]  # 199 false + 1 undef -> unit


def test_get_current_assignment_and_pending_unsets():
    # GpuClauseSharerImpl.cu:238-244 + Assigs.cu:187-192: buffered unsets are overlaid
    sh = make(4, 1)
    sh.addClause(-1, [mkLit(3)])
    assert sh.trySetSolverValues(0, [mkLit(0), mkLit(1, True)])
    assert sh.getCurrentAssignment(0, 4).tolist() == [TRUE, FALSE, UNDEF, UNDEF]
    for i in range(32):
        assert sh.trySendAssignment(0) == i
    assert not sh.trySetSolverValues(0, [mkLit(2)])       # no free slot: nothing changes
    sh.unsetSolverValues(0, [mkLit(0)])                     # buffered
    assert sh.getCurrentAssignment(0, 4).tolist() == [UNDEF, FALSE, UNDEF, UNDEF]
    execute(sh)
    assert sh.trySetSolverValues(0, [mkLit(2)])             # flushes the buffered unset first
    assert sh.getCurrentAssignment(0, 4).tolist() == [UNDEF, FALSE, TRUE, UNDEF]
    assert sh.getOneSolverStat(0, OneSolverStats.failuresToFindAssig) == 1
    assert sh.getLastAssigAllReported(0) == 0
    popped(sh, 0)
    assert sh.getLastAssigAllReported(0) == 32


def test_stats_names_and_counts():
    sh = make(2, 1)
    assert sh.getGlobalStatCount() == 13 and sh.getOneSolverStatCount() == 6
    assert sh.getGlobalStatName(GlobalStats.gpuReports) == "gpuReports"
    assert sh.getOneSolverStatName(OneSolverStats.reportedClausesBinary) == "reportedClausesBinary"
    free, total = sh.getGpuMemInfo()
    assert 0 < free <= total
    assert not sh.hasRunOutOfGpuMemoryOnce()


def test_no_run_on_empty_database():
    # GpuRunner.cu:284-288: with no clause nothing is collected, so slots fill up
    sh = make(2, 1)
    for i in range(32):
        assert sh.trySendAssignment(0) == i
    execute(sh)
    assert sh.trySendAssignment(0) == -1
    assert sh.getGlobalStat(GlobalStats.gpuRuns) == 0
    sh.addClause(-1, [mkLit(0)])
    execute(sh)
    assert sh.trySendAssignment(0) == 32
    assert int(sh.debugLastHits()["mask"][0]) == 0xFFFFFFFF


def test_large_hit_list_parallel_hand_over():
    # > 8192 hits takes the per-solver parallel hand-over path: same batches, same order
    n, nsolvers = 20000, 4
    sh = make(n, nsolvers, gpuBlockCountGuideline=-1, gpuThreadsPerBlockGuideline=-1)
    for i in range(n):
        sh.addClause(-1, [mkLit(i), mkLit((i + 1) % n, True)])
    for s in range(nsolvers):
        # solver s: variable v false when v % (s + 2) == 0, else undefined
        assert sh.trySetSolverValues(s, [mkLit(v, True) for v in range(0, n, s + 2)])
        sh.trySendAssignment(s)
    execute(sh)
    hits = sh.debugLastHits()
    for s in range(nsolvers):
        m = s + 2
        # clause i = (i, not i+1) fires iff i is false (no true literal, at most one undefined)
        # and i+1 is not false (else its negation is true)
        expect = [i for i in range(n) if i % m == 0 and ((i + 1) % n) % m != 0]
        got = hits[hits["solver_id"] == s]["clause_id"].tolist()
        assert got == expect
        pops = popped(sh, s)
        assert [p[1] for p in pops] == expect           # handed over in clause order
        assert pops[0][0] == [mkLit(expect[0]), mkLit((expect[0] + 1) % n, True)]
    assert len(hits) > 8192


def test_more_distinct_clause_lengths_than_the_kernels_cache():
    """The kernels search the length directory from shared memory (128 entries) and read what lies past it from the
    directory itself.  150 distinct lengths, every clause falsified: each solver must get every clause back with its
    own id and literals -- the shortest lengths are the directory's LAST entries (up to round 2 their hits were
    resolved against another length's arena: the import-latency harness lost every binary probe)."""
    nlen, nsolvers = 150, 2
    sh = make(400, nsolvers)
    sh.setMaxClauseLen(200)
    want = {}
    for ln in range(1, nlen + 1):
        for rep in range(2):
            lits = [mkLit((7 * ln + 13 * rep + i) % 397) for i in range(ln)]
            cid = sh.addClause(-1, lits)
            assert cid >= 0
            want[cid] = lits
    for s in range(nsolvers):
        assert sh.trySetSolverValues(s, [mkLit(v, True) for v in range(400)])
        assert sh.trySendAssignment(s) >= 0
    execute(sh)
    hits = sh.debugLastHits()
    assert sorted(hits["clause_id"].tolist()) == sorted(list(want) * nsolvers)
    for s in range(nsolvers):
        got = {}
        while (x := sh.popReportedClause(s)) is not None:
            got[int(x[1])] = [int(l) for l in x[0]]
        assert got == want
