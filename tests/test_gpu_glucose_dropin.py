"""Drop-in proof: the reference's real GPU portfolio solver (glucose-syrup/gpu, unmodified sources
compiled with plain g++ in the container where /root/reference exists) linked against
libgpushare_b200.so through the shim solves CNF instances on the GPU box; the same solver linked
against the reference's own GPU library must give the same verdicts."""
import os
import re
import subprocess

import numpy as np
import pytest

from golden_replay import load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MINE = os.path.join(ROOT, "oracle", "_ref", "glucose-gpu-b200")
REF = os.path.join(ROOT, "oracle", "_ref", "glucose-gpu-ref")

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(MINE), reason="glucose-gpu-b200 not built")]


def random_3sat(path, n, m, seed):
    rng = np.random.default_rng(seed)
    with open(path, "w") as f:
        f.write(f"p cnf {n} {m}\n")
        for _ in range(m):
            vs = rng.choice(n, size=3, replace=False) + 1
            sg = rng.integers(0, 2, size=3) * 2 - 1
            f.write(" ".join(str(int(v * s)) for v, s in zip(vs, sg)) + " 0\n")


def solve(exe, cnf, threads=3, env=None, extra=()):
    r = subprocess.run([exe, f"-thread-count={threads}", "-verb=0", *extra, cnf], capture_output=True, text=True, timeout=300,
                       env=env)
    m = re.search(r"^s (SATISFIABLE|UNSATISFIABLE|INDETERMINATE)", r.stdout, re.M)
    assert m, r.stdout[-2000:] + r.stderr[-2000:]
    # MiniSat convention (satUtils/InitHelper.h:63-65): 10 SAT, 20 UNSAT
    assert r.returncode == {"SATISFIABLE": 10, "UNSATISFIABLE": 20, "INDETERMINATE": 0}[m.group(1)]
    return m.group(1), r.stdout


def check_model(cnf, stdout):
    vals = {}
    for line in stdout.splitlines():
        if line.startswith("v "):
            for t in line[2:].split():
                if t != "0":
                    vals[abs(int(t))] = int(t) > 0
    if not vals:
        return
    for line in open(cnf):
        if line[0] in "pc":
            continue
        lits = [int(t) for t in line.split()[:-1]]
        assert any(vals.get(abs(l), False) == (l > 0) for l in lits), line


def test_real_glucose_gpu_solver_runs_on_our_library(tmp_path):
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
        v, out = solve(MINE, cnf)
        assert v in ("SATISFIABLE", "UNSATISFIABLE")
        if v == "SATISFIABLE":
            check_model(cnf, out)
        verdicts.append(v)
        if os.path.exists(REF):
            assert solve(REF, cnf)[0] == v
    assert "SATISFIABLE" in verdicts and "UNSATISFIABLE" in verdicts
