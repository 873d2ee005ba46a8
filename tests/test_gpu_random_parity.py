"""Random API traffic: the CUDA path (through the C ABI) against the snapshot model built on the
CPU oracle -- sorted (clause, solver, mask) triples must be bit-identical, in production and in
dense mode, for several solver counts, slot wrap-around, full slots and buffered unsets."""
import numpy as np
import pytest

from gpusharesat_b200 import GpuClauseSharer, GpuClauseSharerOptions, mkLit
from oracle_lib import SharerModel

pytestmark = pytest.mark.gpu


def run_random(nvars, nsolvers, nrounds, seed, dense=False, blocks=3, threads=32, max_len=12, report=2,
               clauses_per_round=40, p_false=0.55, p_undef=0.25):
    rng = np.random.default_rng(seed)
    sh = GpuClauseSharer(GpuClauseSharerOptions(gpuBlockCountGuideline=blocks, gpuThreadsPerBlockGuideline=threads,
                                                minGpuLatencyMicros=0, initReportCountPerCategory=report))
    sh.setVarCount(nvars)
    sh.setCpuSolverCount(nsolvers)
    sh.debugSetDense(dense)
    model = SharerModel(nvars, nsolvers)
    total_hits = 0
    for r in range(nrounds):
        for _ in range(int(rng.integers(0, clauses_per_round))):
            n = int(rng.integers(1, max_len + 1))
            lits = [mkLit(int(rng.integers(0, nvars)), bool(rng.integers(0, 2))) for _ in range(n)]
            a, b = sh.addClause(-1, lits), model.addClause(lits)
            assert a == b
        for s in range(nsolvers):
            for _ in range(int(rng.integers(0, 40))):  # sometimes more than 32: slots fill up
                k = int(rng.integers(0, max(2, nvars // 3)))
                vs = rng.choice(nvars, size=min(k, nvars), replace=False)
                x = rng.random(len(vs))
                unset = [mkLit(int(v)) for v, xx in zip(vs, x) if xx < p_undef]
                sets = [mkLit(int(v), bool(xx < p_undef + p_false)) for v, xx in zip(vs, x) if xx >= p_undef]
                sh.unsetSolverValues(s, unset); model.unsetSolverValues(s, unset)
                assert sh.trySetSolverValues(s, sets) == model.trySetSolverValues(s, sets)
                assert sh.trySendAssignment(s) == model.trySendAssignment(s)
            assert sh.getCurrentAssignment(s, nvars).tolist() == np.where(
                np.isin(np.arange(nvars), [l >> 1 for l in model.to_unset[s]]), 2, model.vals[s]).tolist()
        sh.gpuRun()
        sh.gpuRun()
        expect = model.run()
        got = sh.debugLastHits()
        if expect is None:
            assert len(got) == 0
            continue
        # the second gpuRun() started a run with no new assignment; nothing to compare for it
        model.run()
        assert len(got) == len(expect), (r, len(got), len(expect))
        assert np.array_equal(got, expect), r
        total_hits += len(got)
        while sh.popReportedClause(int(rng.integers(0, nsolvers))) is not None:
            pass
    return total_hits


@pytest.mark.parametrize("nsolvers", [1, 2, 3, 5, 32])
def test_random_traffic_matches_oracle(nsolvers):
    assert run_random(40, nsolvers, 12, seed=100 + nsolvers) > 0


@pytest.mark.parametrize("nsolvers", [1, 4, 32])
def test_random_traffic_dense_mode(nsolvers):
    assert run_random(40, nsolvers, 8, seed=200 + nsolvers, dense=True) > 0


def test_more_than_32_solvers():
    # the reference never checks solvers 32.. (Assigs.cu:409-425); here every group of 32 solvers
    # has its own aggregate word
    assert run_random(30, 40, 6, seed=7) > 0
    assert run_random(30, 70, 4, seed=8, dense=True) > 0


def test_default_grid_and_long_clauses():
    assert run_random(300, 8, 6, seed=9, blocks=-1, threads=-1, max_len=100, clauses_per_round=400,
                      p_false=0.9, p_undef=0.08, report=-1) > 0


def test_var_count_growth_rebuilds_tables():
    rng = np.random.default_rng(5)
    sh = GpuClauseSharer(GpuClauseSharerOptions(minGpuLatencyMicros=0))
    model = SharerModel(64, 2)
    sh.setVarCount(16); sh.setCpuSolverCount(2)
    for nv in (16, 40, 64):
        sh.setVarCount(nv)
        for _ in range(50):
            lits = [mkLit(int(rng.integers(0, nv)), bool(rng.integers(0, 2))) for _ in range(int(rng.integers(1, 4)))]
            sh.addClause(-1, lits); model.addClause(lits)
        for s in range(2):
            for _ in range(5):
                vs = rng.choice(nv, size=nv // 2, replace=False)
                sets = [mkLit(int(v), bool(rng.integers(0, 2))) for v in vs]
                assert sh.trySetSolverValues(s, sets) and model.trySetSolverValues(s, sets)
                assert sh.trySendAssignment(s) == model.trySendAssignment(s)
        sh.gpuRun(); sh.gpuRun()
        expect = model.run(); model.run()
        assert np.array_equal(sh.debugLastHits(), expect)
        assert len(expect) > 0
