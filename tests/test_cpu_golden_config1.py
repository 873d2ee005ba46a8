"""CPU tier: BASELINE.json configs[0] -- the golden inputs recorded from the reference's own CPU
solver (glucose-syrup/simp on random 3-SAT n=300 m=1278; tests/golden/make_config1.py) replayed on
the CPU oracle must reproduce the committed expected hits."""
from golden_replay import as_lists, load, model_run, replay
from oracle_lib import SharerModel


def test_oracle_reproduces_config1_golden_hits():
    fx = load()
    assert fx["nvars"] == 300 and fx["cnf"].startswith("p cnf 300 1278")
    model = SharerModel(fx["nvars"], fx["nsolvers"])
    runs = replay(fx["events"], model, model_run, is_model=True)
    assert len(runs) == len(fx["expected_hits_per_run"]) == 230
    assert as_lists(runs) == fx["expected_hits_per_run"]
    assert sum(len(r) for r in runs) > 1000
    # learned clauses of a real CDCL run: every recorded clause is non-empty, literals within range
    n_cl = 0
    for ev in fx["events"]:
        if ev[0] == "c":
            n_cl += 1
            assert len(ev) >= 2 and all(0 <= l < 600 for l in ev[1:])
    assert n_cl == 1500
