"""Several devices behind ONE sharer in ONE process (GPUSHARE_DEVICES=N, csrc/multi.cu), one rank per
DISTINCT device: the union of the devices' hits must equal the single-device sharer's and the CPU
oracle's, every solver must be handed the same clauses, reduceDb must leave the same database, and
the reference's real GPU portfolio solver must solve on it through the unmodified header.  Needs at
least two GPUs (gpurun --gpus 2); skipped on a one-GPU box."""
import os
import subprocess
import sys

import numpy as np
import pytest

from gpusharesat_b200 import GpuClauseSharer, GpuClauseSharerOptions, mkLit
from oracle_lib import SharerModel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


pytestmark = [pytest.mark.gpu, pytest.mark.skipif(_gpus() < 2, reason="needs two CUDA devices")]


def make(devices, **opts):
    if devices > 1:
        os.environ["GPUSHARE_DEVICES"] = str(devices)
    try:
        return GpuClauseSharer(GpuClauseSharerOptions(minGpuLatencyMicros=0, **opts))
    finally:
        os.environ.pop("GPUSHARE_DEVICES", None)


def model_clauses(model):
    return len(model.clauses) if hasattr(model, "clauses") else 1 << 60


def pop_all(sh, s):
    out = []
    while True:
        r = sh.popReportedClause(s)
        if r is None:
            return out
        out.append((r[1], tuple(r[0])))


@pytest.mark.parametrize("devices,nsolvers", [(2, 3), (2, 40), (min(4, max(2, _gpus())), 8)])
def test_multi_device_equals_single_device_and_oracle(devices, nsolvers):
    rng = np.random.default_rng(40 + devices + nsolvers)
    nvars = 80
    a, b = make(devices), make(1)
    model = SharerModel(nvars, nsolvers)
    for sh in (a, b):
        sh.setVarCount(nvars)
        sh.setCpuSolverCount(nsolvers)
    total = 0
    for r in range(8):
        for _ in range(700):  # several 128-clause tiles per length, so that every device has a share
            n = int(rng.integers(1, 6))
            lits = [mkLit(int(rng.integers(0, nvars)), bool(rng.integers(0, 2))) for _ in range(n)]
            s = int(rng.integers(-1, nsolvers))
            assert len({sh.addClause(s, lits) for sh in (a, b)}) == 1
            model.addClause(lits)
        for s in range(nsolvers):
            for _ in range(int(rng.integers(0, 10))):
                vs = rng.choice(nvars, size=int(rng.integers(1, nvars // 2)), replace=False)
                x = rng.random(len(vs))
                unset = [mkLit(int(v)) for v, xx in zip(vs, x) if xx < 0.2]
                sets = [mkLit(int(v), bool(xx < 0.8)) for v, xx in zip(vs, x) if xx >= 0.2]
                for sh in (a, b, model):
                    sh.unsetSolverValues(s, unset)
                assert len({sh.trySetSolverValues(s, sets) for sh in (a, b, model)}) == 1
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
            # every solver's batch merges the devices' slices back into the single-device order
            assert pop_all(a, s) == pop_all(b, s), (r, s)
            assert a.getLastAssigAllReported(s) == b.getLastAssigAllReported(s)
    assert total > 0
    # reduceDb: the activities live on device 0, every device must compact identically (the oracle model
    # keeps every clause, so from here on the two sharers are compared with each other only)
    for r in range(3):
        for sh in (a, b):
            sh.reduceDb()
        assert a.getGlobalStat(0) == b.getGlobalStat(0) and a.getGlobalStat(1) == b.getGlobalStat(1)
        assert a.getGlobalStat(0) < model_clauses(model) if r == 0 else True
        for s in range(nsolvers):
            vs = rng.choice(nvars, size=nvars // 2, replace=False)
            sets = [mkLit(int(v), bool(rng.random() < 0.8)) for v in vs]
            for sh in (a, b):
                sh.unsetSolverValues(s, [mkLit(v) for v in range(nvars)])
                assert sh.trySetSolverValues(s, sets)
                assert sh.trySendAssignment(s) >= 0
        for sh in (a, b):
            sh.gpuRun()
            sh.gpuRun()
        ha, hb = a.debugLastHits(), b.debugLastHits()
        assert len(ha) > 0 and np.array_equal(ha, hb), r
        for s in range(nsolvers):
            assert pop_all(a, s) == pop_all(b, s), (r, s)
        for _ in range(300):
            n = int(rng.integers(3, 6))
            lits = [mkLit(int(rng.integers(0, nvars)), bool(rng.integers(0, 2))) for _ in range(n)]
            assert len({sh.addClause(-1, lits) for sh in (a, b)}) == 1
    for st in (0, 1, 2, 3, 4, 5, 6, 7, 8):
        assert a.getGlobalStat(st) == b.getGlobalStat(st), st


def test_glucose_portfolio_solver_on_two_devices(tmp_path):
    """the reference's real GPU solver (unmodified sources, linked through the shim) with
    GPUSHARE_DEVICES=2: the factory keeps its signature, the device count comes from the environment"""
    from test_gpu_glucose_dropin import MINE, check_model, random_3sat, solve
    if not os.path.exists(MINE):
        pytest.skip("glucose-gpu-b200 not built")
    for i, (n, m) in enumerate([(250, 1000), (200, 1000), (150, 900)]):
        p = str(tmp_path / f"r{i}.cnf")
        random_3sat(p, n, m, 100 + i)
        v, out = solve(MINE, p, env=dict(os.environ, GPUSHARE_DEVICES="2"), extra=("-model",))
        assert v in ("SATISFIABLE", "UNSATISFIABLE")
        if v == "SATISFIABLE":
            check_model(p, out)
        assert solve(MINE, p)[0] == v
