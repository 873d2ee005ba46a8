"""bench.py's JSON contract, as far as it can be checked without a GPU: the reference arm (CPU port of
the reference algorithm, oracle/) runs here, and the committed B200 lines under profiles/ carry every
key the contract names."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e,
                         timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_line():
    small = ["--impl", "reference", "--steps", "2", "--warmup", "1", "--clauses", "60000", "--vars", "20000"]
    (line,) = _run(small)
    assert BASE_KEYS <= set(line) and line["impl"] == "reference"
    assert line["metric"] == "clause_literal_x_assignment_checks_per_sec" and line["unit"] == "checks/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["dtype"] == "u32"
    assert line["value"] > 0 and line["steps"] == 2 and line["warmup"] == 1
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]
    # under torchrun only rank 0 works and prints; weak scaling names the N-fold database
    assert _run(small + ["--gpus", "2"], {"RANK": "1", "WORLD_SIZE": "2"}) == []
    (two,) = _run(small + ["--gpus", "2"], {"RANK": "0", "WORLD_SIZE": "2"})
    assert two["n_gpus"] == 2 and two["scaling"] == "weak" and two["config"]["clauses"] == 120000
    assert two["config"]["clauses_per_gpu"] == 60000


def test_committed_b200_lines_carry_the_contract_keys():
    line = json.loads(open(os.path.join(ROOT, "profiles", "r01_bench.json")).read().strip().splitlines()[-1])
    assert BASE_KEYS | {"gpu_launches", "clocks", "roofline", "cpu_baseline"} <= set(line)
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"])
    assert line["roofline"]["bound"] == "hbm" and line["roofline"]["unit"] == "GB/s"
    assert abs(line["roofline"]["frac"] - line["roofline"]["achieved"] / line["roofline"]["peak"]) < 1e-9
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["gpu_launches"] > 0 and line["parity_sample"]["identical"] is True
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    scale = json.load(open(os.path.join(ROOT, "profiles", "r01_scale.json")))
    assert scale["runs"]["8_weak_peer"]["n_gpus"] == 8 and scale["runs"]["8_weak_peer"]["scaling"] == "weak"
