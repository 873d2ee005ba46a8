#!/usr/bin/env python
"""Generates tests/golden/config1.json.gz (BASELINE.json configs[0]).

Needs /root/reference: builds oracle/_ref/golden_recorder (the reference's CPU-only glucose-syrup
solver compiled in place + oracle/golden_recorder.cc), runs it on random 3-SAT n=300 m=1278 and
stores (a) the recorded call sequence -- learned clauses and trail snapshots exactly as a
GPU-helped solver thread would send them -- and (b) the hit triples the CPU oracle's snapshot model
produces for every run.  The fixture travels to the GPU box; /root/reference does not."""
import gzip
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "recorder"])
    ev_path, cnf_path = "/tmp/config1_events.txt", "/tmp/config1.cnf"
    out = subprocess.check_output([os.path.join(ROOT, "oracle", "_ref", "golden_recorder"), "300", "1278", "1500",
                                   ev_path, cnf_path], text=True)
    events = []
    for line in open(ev_path):
        parts = line.split()
        events.append([parts[0]] + [int(x) for x in parts[1:]])
    from golden_replay import as_lists, model_run, replay
    from oracle_lib import SharerModel
    model = SharerModel(300, 1)
    runs = replay(events, model, model_run, is_model=True)
    fixture = {
        "source": "glucose-syrup/simp (reference CPU solver) on random 3-SAT n=300 m=1278, hooks core/Solver.h:265-268",
        "recorder_output": out.strip().splitlines()[-1],
        "nvars": 300, "nsolvers": 1,
        "cnf": open(cnf_path).read(),
        "events": events,
        "expected_hits_per_run": as_lists(runs),
    }
    with gzip.open(os.path.join(HERE, "config1.json.gz"), "wt") as f:
        json.dump(fixture, f, separators=(",", ":"))
    n = sum(len(r) for r in runs)
    print(f"{len(events)} events, {len(runs)} runs, {n} hit records -> config1.json.gz")


if __name__ == "__main__":
    main()
