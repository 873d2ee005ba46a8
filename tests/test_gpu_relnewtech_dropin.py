"""Drop-in proof for the second consumer of the boundary: rel-newtech's GPU build
(rel-newtech/gpu/Main.cc:57,98, core/Solver.cc:2891-2978, gpu/GpuMultiSolver.cc:115-131; unmodified
sources compiled with plain g++ where /root/reference exists: oracle/Makefile target `relnewtech`)
linked against libgpushare_b200.so through the shim solves CNF instances on the GPU box with the
verdicts of the same solver linked against the reference's own GPU library."""
import os

import pytest

from golden_replay import load
from test_gpu_glucose_dropin import check_model, random_3sat, solve

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MINE = os.path.join(ROOT, "oracle", "_ref", "rel-newtech-gpu-b200")
REF = os.path.join(ROOT, "oracle", "_ref", "rel-newtech-gpu-ref")

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(MINE), reason="rel-newtech-gpu-b200 not built")]


def test_real_rel_newtech_gpu_solver_runs_on_our_library(tmp_path):
    cases = []
    p = str(tmp_path / "config1.cnf")
    open(p, "w").write(load()["cnf"])
    cases.append(p)
    for i, (n, m) in enumerate([(250, 1000), (200, 1000), (150, 900)]):  # under / over the threshold
        p = str(tmp_path / f"r{i}.cnf")
        random_3sat(p, n, m, 100 + i)
        cases.append(p)
    verdicts = []
    for cnf in cases:
        v, out = solve(MINE, cnf, extra=("-model",))
        assert v in ("SATISFIABLE", "UNSATISFIABLE")
        if v == "SATISFIABLE":
            check_model(cnf, out)
        verdicts.append(v)
        if os.path.exists(REF):
            assert solve(REF, cnf)[0] == v
    assert "SATISFIABLE" in verdicts and "UNSATISFIABLE" in verdicts


def test_rel_newtech_with_64_solver_threads_and_lifted_clause_length(tmp_path):
    """BASELINE config 5's shape through the UNMODIFIED header: 64 solver threads (the reference never
    checks solvers >= 32, Assigs.cu:409-425) and GPUSHARE_MAX_CLAUSE_LEN=200 (reference: 100)."""
    p = str(tmp_path / "r.cnf")
    random_3sat(p, 220, 935, 7)
    env = dict(os.environ, GPUSHARE_MAX_CLAUSE_LEN="200")
    v, out = solve(MINE, p, threads=64, env=env, extra=("-model",))
    assert v in ("SATISFIABLE", "UNSATISFIABLE")
    if v == "SATISFIABLE":
        check_model(p, out)
