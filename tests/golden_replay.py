"""Replay of the config-1 golden event log (recorded from the reference's CPU solver by
oracle/golden_recorder.cc) against anything with the GpuClauseSharer method names."""
import gzip
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config1.json.gz")


def load():
    with gzip.open(GOLDEN, "rt") as f:
        return json.load(f)


def replay(events, sh, hits_of_last_run, solver=0, is_model=False):
    """returns the list of per-run hit arrays [(clause_id, solver, mask), ...]"""
    runs = []
    for ev in events:
        tag, lits = ev[0], ev[1:]
        if tag == "c":
            # GpuHelpedSolver.cc:94-101: the learner exports with its own solver id
            if is_model:
                sh.addClause(lits)
            else:
                sh.addClause(solver, lits)
        elif tag == "u":
            sh.unsetSolverValues(solver, lits)
        elif tag == "s":
            assert sh.trySetSolverValues(solver, lits)
        elif tag == "a":
            assert sh.trySendAssignment(solver) >= 0
        elif tag == "r":
            runs.append(hits_of_last_run(sh))
    return runs


def model_run(m):
    h = m.run()
    m.run()
    return h if h is not None else np.zeros(0, dtype=[("clause_id", "<i8"), ("solver_id", "<i4"), ("mask", "<u4")])


def lib_run(sh):
    sh.gpuRun()
    sh.gpuRun()
    h = sh.debugLastHits()
    while sh.popReportedClause(0) is not None:
        pass
    return h


def as_lists(runs):
    return [[[int(a), int(b), int(c)] for a, b, c in r.tolist()] for r in runs]
