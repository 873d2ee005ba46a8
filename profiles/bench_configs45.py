#!/usr/bin/env python
"""Measurements for BASELINE.json configs[3] and configs[4] (the "next" rows of SURVEY.md 8f):

config 4  streaming clause ingest + reduceDb at 8 M-clause scale: 32 solver threads' worth of
          assignments per run while clauses keep arriving; per-run latency as the database grows,
          sustained ingest rate, reduceDb time.
config 5  64 solver threads (two aggregate groups), long-clause-heavy database (5 % of the clauses
          have 101-200 literals): end-to-end import latency = trySendAssignment -> the clause is
          handed back by popReportedClause, with a GPU thread spinning on gpuRun().

usage: python profiles/bench_configs45.py > profiles/r01_configs45.json   (on the GPU box)"""
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import synth  # noqa: E402
from gpusharesat_b200 import GpuClauseSharer, GpuClauseSharerOptions, GlobalStats, mkLit  # noqa: E402


def config4(total=8_000_000, nvars=2_000_000, solvers=32, chunk=100_000):
    sig = synth.sigma(nvars, 21)
    offsets, lits = synth.clauses(total, nvars, 30, sig, 0.98, 22)
    sh = GpuClauseSharer(GpuClauseSharerOptions(minGpuLatencyMicros=0))
    sh.setVarCount(nvars)
    sh.setCpuSolverCount(solvers)
    streams = [synth.Stream(nvars, sig, 0.01, 0.0005, 500 + s) for s in range(solvers)]
    for s, st in enumerate(streams):  # initial trail
        sets, unsets = st.next()
        assert sh.trySetSolverValues(s, sets)
        sh.trySendAssignment(s)
    rows, t_ingest = [], 0.0
    added = 0
    while added < total:
        hi = min(total, added + chunk)
        off = offsets[added:hi + 1] - offsets[added]
        t0 = time.perf_counter()
        sh.addClausesBulk(off, lits[offsets[added]:offsets[hi]])
        for s, st in enumerate(streams):  # two fresh assignments per solver and run
            for _ in range(2):
                sets, unsets = st.next()
                sh.unsetSolverValues(s, unsets)
                sh.trySetSolverValues(s, sets)
                sh.trySendAssignment(s)
        t1 = time.perf_counter()
        sh.gpuRun()
        t2 = time.perf_counter()
        for s in range(solvers):
            while sh.popReportedClause(s) is not None:
                pass
        added = hi
        t_ingest += t2 - t0
        if (added // chunk) % 10 == 0 or added == total:
            ph = sh.debugLastRunTimes()
            rows.append({"clauses": added, "gpuRun_ms": (t2 - t1) * 1e3, "device_check_us": ph[2] if ph else None,
                         "h2d_us": ph[0] if ph else None})
    sh.gpuRun()
    t0 = time.perf_counter()
    sh.reduceDb()
    t_reduce = time.perf_counter() - t0
    after = sh.getGlobalStat(GlobalStats.gpuClauses)
    t0 = time.perf_counter()
    sh.gpuRun(); sh.gpuRun()
    t_after = time.perf_counter() - t0
    return {"workload": f"{total} clauses (Luby mix) over {nvars} vars arriving {chunk} per run, {solvers} solvers x 2 assignments per run",
            "ingest_clauses_per_s_incl_run": total / t_ingest, "growth": rows,
            "reduce_db_ms": t_reduce * 1e3, "clauses_after_reduce": after, "first_two_runs_after_reduce_ms": t_after * 1e3,
            "out_of_memory": sh.hasRunOutOfGpuMemoryOnce()}


def config5(nclauses=2_000_000, nvars=500_000, solvers=64, probes=200):
    rng = np.random.default_rng(31)
    sig = synth.sigma(nvars, 31)
    offsets, lits = synth.clauses(nclauses, nvars, 30, sig, 0.98, 32)
    # 5 % long clauses, 101..200 literals, satisfied by sigma with high probability
    nlong = nclauses // 20
    lens = rng.integers(101, 201, size=nlong)
    loff = np.zeros(nlong + 1, dtype=np.int64)
    np.cumsum(lens, out=loff[1:])
    lv = rng.integers(0, nvars, size=loff[-1])
    agree = rng.random(loff[-1]) < 0.98
    lsign = np.where(agree, sig[lv], 1 - sig[lv])
    llits = (2 * lv + lsign).astype(np.int32)
    sh = GpuClauseSharer(GpuClauseSharerOptions(minGpuLatencyMicros=0))
    sh.setMaxClauseLen(200)
    sh.setVarCount(nvars)
    sh.setCpuSolverCount(solvers)
    sh.addClausesBulk(offsets, lits)
    sh.addClausesBulk(loff, llits)
    # probe clauses over private variables: unit under "all their variables false but the last"
    probe_vars = nvars - 1 - np.arange(probes * 2)
    probe_ids = []
    for k in range(probes):
        a, b = int(probe_vars[2 * k]), int(probe_vars[2 * k + 1])
        probe_ids.append(sh.addClause(-1, [mkLit(a), mkLit(b)]))
    for s in range(solvers):
        v = np.arange(nvars - probes * 2)
        keep = rng.random(v.size) > 0.01
        sets = (2 * v[keep] + sig[v[keep]]).astype(np.int32)
        assert sh.trySetSolverValues(s, sets)
        sh.trySendAssignment(s)
    stop = threading.Event()

    def gpu_thread():
        while not stop.is_set():
            sh.gpuRun()

    th = threading.Thread(target=gpu_thread)
    th.start()
    time.sleep(0.5)
    for s in range(solvers):
        while sh.popReportedClause(s) is not None:
            pass
    lat = []
    for k in range(probes):
        s = k % solvers
        a = int(probe_vars[2 * k])
        t0 = time.perf_counter()
        assert sh.trySetSolverValues(s, [mkLit(a, True)])   # a false, b undefined -> the probe clause is unit
        sh.trySendAssignment(s)
        got = None
        while got is None or got[1] != probe_ids[k]:
            got = sh.popReportedClause(s)
            if time.perf_counter() - t0 > 1:
                break
        lat.append(time.perf_counter() - t0)
        sh.unsetSolverValues(s, [mkLit(a)])
    stop.set()
    th.join()
    lat = np.array(lat) * 1e6
    runs = sh.getGlobalStat(GlobalStats.gpuRuns)
    return {"workload": f"{nclauses} clauses (Luby mix) + {nlong} long clauses (101-200 lits) over {nvars} vars, {solvers} solver threads",
            "import_latency_us": {"p50": float(np.percentile(lat, 50)), "p90": float(np.percentile(lat, 90)),
                                  "p99": float(np.percentile(lat, 99)), "max": float(lat.max())},
            "probes": probes, "gpu_runs": runs, "timeouts": int(np.sum(lat > 0.99e6))}


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "both"
    out = {}
    if which in ("both", "4"):
        out["config4_streaming_ingest"] = config4()
    if which in ("both", "5"):
        out["config5_import_latency"] = config5()
    print(json.dumps(out, indent=1))
