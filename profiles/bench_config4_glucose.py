#!/usr/bin/env python
"""BASELINE configs[3] as stated: the reference's REAL GPU portfolio solver (glucose-syrup/gpu, unmodified
sources) with 32 solver threads on a synthetic SAT-competition-shaped instance (2 M variables, 8 M clauses:
planted satisfiable, 88 % ternary clauses at a clause / variable ratio of 4, the rest of length 2 and 4-9), the learned clauses streaming into
the GPU clause database -- once linked against libgpushare_b200.so through the shim (glucose-gpu-b200) and
once against the reference's own GPU library recompiled for sm_100a (glucose-gpu-ref), for the same wall
time.  Prints one JSON object with the last periodic statistics of both (GPU runs, clause tests, reports,
clauses on the GPU, reduceDbs, imports per solver).

usage (GPU box): python profiles/bench_config4_glucose.py [--seconds 75] > gpurun_out/config4_glucose.json"""
import argparse
import ctypes as C
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402


def write_cnf(path, nvars, nclauses, seed):
    L = synth.load_library()
    L.gss_synth_write_cnf.restype = C.c_int
    L.gss_synth_write_cnf.argtypes = [C.c_char_p, C.c_int, C.c_int64, C.POINTER(C.c_double), C.c_int, C.c_double, C.c_uint64]
    w = [0.02, 0.88, 0.04, 0.03, 0.01, 0.01, 0.005, 0.005]  # lengths 2..9: 3-SAT dominated, clause / variable ratio 4
    arr = (C.c_double * len(w))(*w)
    assert L.gss_synth_write_cnf(path.encode(), nvars, nclauses, arr, len(w), 0.5, seed) == 0


def last_stats(text):
    """the last {"type": "periodicStats" ...} object the solver printed (every line carries the comment prefix "c")"""
    text = re.sub(r"^c ?", "", text, flags=re.M)
    best = None
    for m in re.finditer(r'\{\s*"type"\s*:\s*"periodicStats"', text):
        depth, i = 0, m.start()
        for j in range(i, len(text)):
            if text[j] == "{":
                depth += 1
            elif text[j] == "}":
                depth -= 1
                if depth == 0:
                    try:
                        best = json.loads(text[i:j + 1])
                    except Exception:
                        pass
                    break
    return best


def run(exe, cnf, threads, seconds, env=None):
    t0 = time.time()
    p = subprocess.run(["timeout", "-s", "INT", str(seconds), exe, f"-thread-count={threads}", "-verb=1",
                        "-write-stats-period-sec=5", "-max-memory=40000", "-mem-lim=60000", "-no-pre", cnf],
                       capture_output=True, text=True, env=env)
    wall = time.time() - t0
    verdict = re.search(r"^s (\w+)", p.stdout, re.M)
    st = last_stats(p.stdout)
    out = {"wall_s": wall, "verdict": verdict.group(1) if verdict else None, "rc": p.returncode}
    if st:
        g = st.get("globalStats", {})
        out["realTime_of_last_stats"] = st.get("realTime")
        out["global"] = g
        sol = st.get("solverStats", [])
        keys = ("conflicts", "nbImported", "nbimported", "reportedClauses", "nbexported", "nbexportedunit")
        agg = {}
        for s in sol:
            for k, v in s.items():
                if isinstance(v, (int, float)) and any(x in k.lower() for x in ("import", "export", "conflict", "report", "propag")):
                    agg[k] = agg.get(k, 0) + v
        out["solver_sums"] = agg
        rt = st.get("realTime") or wall
        if g.get("gpuRuns") is not None:
            out["gpu_runs_per_s"] = g["gpuRuns"] / rt
    else:
        out["tail"] = p.stdout[-1500:] + p.stderr[-1500:]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vars", type=int, default=2_000_000)
    ap.add_argument("--clauses", type=int, default=8_000_000)
    ap.add_argument("--threads", type=int, default=32)
    ap.add_argument("--seconds", type=int, default=110)
    ap.add_argument("--cnf", default="/tmp/gss_config4.cnf")
    a = ap.parse_args()
    t0 = time.time()
    write_cnf(a.cnf, a.vars, a.clauses, 41)
    res = {"config": "BASELINE configs[3]", "instance": f"{a.vars} vars, {a.clauses} clauses, planted, 88 % ternary, lengths 2-9",
           "threads": a.threads, "seconds": a.seconds, "host_cores": os.cpu_count(), "cnf_write_s": time.time() - t0}
    for name in ("glucose-gpu-b200", "glucose-gpu-ref"):
        exe = os.path.join(ROOT, "oracle", "_ref", name)
        if os.path.exists(exe):
            res[name] = run(exe, a.cnf, a.threads, a.seconds)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
