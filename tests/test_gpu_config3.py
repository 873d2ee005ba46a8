"""BASELINE.json configs[2] -- the configuration the metric is quoted on -- as a parity test:
10 M clauses (Luby-like lengths 2-30) x 1024 assignments (32 solvers x 32 slots) over 1 M variables,
the SAME seeds bench.py uses.  Three consecutive batches (so the delta path, the deferred collapse and
slot reuse are covered at full size) are compared
  * with the CPU oracle on ALL 10 M clauses (sorted (clause, solver, mask) triples), and
  * with the reference's own GPU checker (oracle/_ref, the unmodified gpuShareLib built for sm_100a)
    driven with the same calls, in the style of the reference's at-scale check
    (glucose-syrup/perftest/perfTest.cu:153-207).
The third batch runs the bench-only dense kernel: identical hit set."""
import os

import numpy as np
import pytest

import ref_lib
import synth
from gpusharesat_b200 import GpuClauseSharer, GpuClauseSharerOptions, GlobalStats
from oracle_lib import check_db

pytestmark = pytest.mark.gpu

NCLAUSES, NVARS, NSOLVERS, NSLOTS = 10_000_000, 1_000_000, 32, 32


def _push_batch(sharers, streams, d, t):
    """every solver exports NSLOTS assignments into every sharer; d/t receive the slot words"""
    d[:] = 0
    t[:] = 0
    for s, st in enumerate(streams):
        for p in range(NSLOTS):
            sets, unsets = st.next()
            for sh in sharers:
                sh.unsetSolverValues(s, unsets)
                assert sh.trySetSolverValues(s, sets)
                assert sh.trySendAssignment(s) >= 0
            v = st.values()
            bit = np.uint32(1 << p)
            d[s][v != 2] |= bit
            t[s][v == 0] |= bit


def test_config3_two_batches_identical_to_oracle_and_reference_gpu():
    sig = synth.sigma(NVARS, 11)
    offsets, lits = synth.clauses(NCLAUSES, NVARS, 30, sig, 0.98, 12)
    sharers = [GpuClauseSharer(GpuClauseSharerOptions(minGpuLatencyMicros=0, verbosity=0))]
    have_ref = ref_lib.available()
    if have_ref:
        # 4000 records per reporter category x 296 categories: the reference cannot drop a hit
        sharers.append(ref_lib.RefSharer(report=4000))
    for sh in sharers:
        sh.setVarCount(NVARS)
        sh.setCpuSolverCount(NSOLVERS)
        assert sh.addClausesBulk(offsets, lits) == 0
    streams = [synth.Stream(NVARS, sig, 0.01, 0.01, 1000 + s) for s in range(NSOLVERS)]
    d = np.zeros((NSOLVERS, NVARS), dtype=np.uint32)
    t = np.zeros((NSOLVERS, NVARS), dtype=np.uint32)
    start = np.full(NSOLVERS, 0xFFFFFFFF, dtype=np.uint32)
    mine = sharers[0]
    for batch in range(3):
        if batch == 2:
            # third batch: the bench-only dense kernel (no filter, no early exit), ours only
            mine.debugSetDense(True)
            sharers = sharers[:1]
        _push_batch(sharers, streams, d, t)
        for sh in sharers:
            sh.gpuRun()
            sh.gpuRun()
        want = check_db(offsets, lits, d, t, start, use_filter=1, nthreads=os.cpu_count() or 8, cap=1 << 19)
        got = mine.debugLastHits()
        assert len(want) > 50_000, len(want)
        assert np.array_equal(got, want), (batch, len(got), len(want))
        if have_ref and batch < 2:
            theirs = sharers[1].debugLastHits()
            assert np.array_equal(theirs, want), (batch, len(theirs), len(want))
        for sh in sharers:
            for s in range(NSOLVERS):
                while sh.popReportedClause(s) is not None:
                    pass
    assert mine.getGlobalStat(GlobalStats.gpuClauses) == NCLAUSES
