#!/usr/bin/env python
"""bench.py -- clause-literal x assignment checks per second on the BASELINE.json workload.

Workload (configs[2], the one the metric is quoted on; it fits one GPU): 10 M clauses with the
Luby-like length mix (2..30, mean 4.41) over 1 M variables, 32 solvers x 32 slots = 1024
assignments per step, "realistic" values from a planted assignment.  One STEP =
one batch of 1024 fresh assignments pushed through the path: delta upload, table update, check
of every clause against every assignment, hits back on the host.  checks per step = L x A
(L = literals in the database), nominal: early exits do not reduce the count.

Numbers on the JSON line
  value      L*A*steps / device time of the steps (CUDA events on the library's stream, max over ranks): everything
             the device does for a batch -- k_apply_direct (which pulls the deltas over PCIe), collapse, k_filter,
             k_exact, k_emit_* (which write the results into page-locked host memory)
  e2e        same metric through the C ABI with HOST buffers: timed region = gss_gpu_run() x 2 per batch
             (collect, launches, wait, activity bumps, hand-over to the solver queues); ms_every_step lists the steps
  roofline   the contract's figure: dense mode (no filter, no early exit: every (literal, 32-slot
             word) pair evaluated) against the slower of HBM bytes and LOP3 issue; LOP3 peak measured in this run,
             HBM peak from MEASURED_PEAKS.json, DRAM traffic from the ncu capture of this command
             (profiles/capture_traffic.py); gather_view = the same sweep against the measured L2 gather ceiling
  roofline_k_filter  the dominant production kernel against HBM (nominal and moved bytes)
  device_step_complete  first event to last event of a step
  cpu_baseline  the reference algorithm (two-level filter, 32 slots bit-parallel) ported to C (oracle/), all host
             cores, on a bounded sample of the same clause database; parity_sample = first batch == that port
  reference_gpu  the reference's own GPU library recompiled for sm_100a (oracle/_ref), same inputs and call pattern
  streamed_db / import_latency  SURVEY 8(f1) / 8(f2)
  host_during_timed_region / remeasured  involuntary context switches of the timing thread inside the timed calls
             and stolen vCPU time; a disturbed region (or a throttled GPU) is measured once more, both attempts kept

--impl reference runs the CPU port as its own arm (the reference has no CPU implementation of this path;
SURVEY.md 8d).  --impl reference-gpu times the reference's own GPU library on the same inputs.  N > 1 (torchrun):
one process per GPU, --scaling weak (default) on the line and strong as a sub-object (or the other way round),
parity_sample in both.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# measured on this pool's B200 (profiles/probes, profiles/r02_probes.jsonl): 64-byte rows gathered at random from a
# 64 MB table (L2 resident), 8 blocks x 256 threads per SM, useful bytes per second
L2_GATHER_GBS = 10223.0
METRIC = "clause_literal_x_assignment_checks_per_sec"
UNIT = "checks/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-gpu"])
    p.add_argument("--clauses", type=int, default=10_000_000)
    p.add_argument("--vars", type=int, default=1_000_000)
    p.add_argument("--solvers", type=int, default=32)
    p.add_argument("--slots", type=int, default=32)
    p.add_argument("--max-len", type=int, default=30)
    p.add_argument("--churn", type=float, default=0.01,
                   help="fraction of variables whose status is re-drawn between consecutive assignments of a solver")
    p.add_argument("--p-undef", type=float, default=0.01)
    p.add_argument("--p-agree", type=float, default=0.98)
    p.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--no-dense", action="store_true")
    p.add_argument("--no-streamed", action="store_true", help="skip the streamed-database pass (streamed_db object)")
    p.add_argument("--stream-chunk", type=int, default=100_000, help="clauses per run of the streamed-database pass")
    p.add_argument("--no-latency", action="store_true", help="skip the import-latency harness (import_latency object)")
    p.add_argument("--no-ref-gpu", action="store_true", help="skip the reference GPU library leg (reference_gpu object)")
    p.add_argument("--ref-gpu-steps", type=int, default=5)
    p.add_argument("--dense-iters", type=int, default=3)
    p.add_argument("--prod-iters", type=int, default=20)
    p.add_argument("--filter-sweep", action="store_true", help="time every variant of the level-1 kernel")
    p.add_argument("--devices", type=int, default=1,
                   help="single process, GPUSHARE_DEVICES=N: the sharer itself shards over N devices (csrc/multi.cu)")
    p.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                   help="N > 1: weak = --clauses per GPU (database grows with N), strong = --clauses in total")
    p.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                   help="N > 1: peer memory windows over NVLink (default) or NCCL broadcast + all-gather")
    p.add_argument("--no-other-scaling", action="store_true", help="N > 1: skip the second pass with the other scaling mode")
    p.add_argument("--parity-clauses", type=int, default=400_000, help="N > 1: clauses of the CPU parity sample")
    p.add_argument("--slot-hits", type=int, default=1 << 21, help="peer exchange: hit records per rank and batch")
    return p.parse_args()


# ------------------------------------------------------------------------------------------------
# inputs
# ------------------------------------------------------------------------------------------------

def make_inputs(a, with_clauses=True):
    import synth
    sig = synth.sigma(a.vars, 11)
    offsets = lits = None
    if with_clauses:
        offsets, lits = synth.clauses(a.clauses, a.vars, a.max_len, sig, a.p_agree, 12)
    return sig, offsets, lits


def make_streams(a, sig):
    import synth
    return [synth.Stream(a.vars, sig, a.p_undef, a.churn, 1000 + s) for s in range(a.solvers)]


def push_batch(sh, streams, slots, pool):
    """what the solver threads do between two GPU runs: every solver exports `slots` assignments
    (unset / set / send per assignment).  Runs on one thread per solver, outside the timed region."""
    def one(s):
        st = streams[s]
        for _ in range(slots):
            sets, unsets = st.next()
            sh.unsetSolverValues(s, unsets)
            if not sh.trySetSolverValues(s, sets):
                raise RuntimeError("no free assignment slot")
            if sh.trySendAssignment(s) < 0:
                raise RuntimeError("no free assignment slot")
    list(pool.map(one, range(len(streams))))


def batch_words(a, sig, first_batch_only=True):
    """def/tru words [solvers][vars] of the FIRST batch of the streams (identical seeds), for the
    CPU arm and the sampled full-size parity check"""
    import synth
    streams = make_streams(a, sig)
    d = np.zeros((a.solvers, a.vars), dtype=np.uint32)
    t = np.zeros((a.solvers, a.vars), dtype=np.uint32)
    for s, st in enumerate(streams):
        for p in range(a.slots):
            st.next()
            v = st.values()
            bit = np.uint32(1 << p)
            d[s][v != 2] |= bit
            t[s][v == 0] |= bit
    start = np.full(a.solvers, (1 << a.slots) - 1 if a.slots < 32 else 0xFFFFFFFF, dtype=np.uint32)
    return d, t, start


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        if os.environ.get("GSS_NO_CLOCK_SAMPLER"):
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


class HostWatch:
    """What happened to the HOST during a timed region: involuntary context switches of the timing thread
    (it was descheduled), vCPU time stolen by the hypervisor (/proc/stat), load average."""

    @staticmethod
    def _steal():
        try:
            f = open("/proc/stat").readline().split()
            return float(f[8]), float(sum(map(float, f[1:9])))
        except Exception:
            return 0.0, 0.0

    def __init__(self):
        import resource
        self.res = resource
        self.who = getattr(resource, "RUSAGE_THREAD", resource.RUSAGE_SELF)
        self.st0 = self._steal()
        self.nivcsw = 0
        self._in = 0

    # the timed calls only: what happens to this thread while the solver threads prepare the next batch does not count
    def enter(self):
        self._in = self.res.getrusage(self.who).ru_nivcsw

    def leave(self):
        self.nivcsw += self.res.getrusage(self.who).ru_nivcsw - self._in

    def stop(self, timed_seconds):
        st = self._steal()
        nivcsw = self.nivcsw
        dt = st[1] - self.st0[1]
        steal = (st[0] - self.st0[0]) / dt if dt > 0 else 0.0
        try:
            load = float(open("/proc/loadavg").read().split()[0])
        except Exception:
            load = None
        out = {"involuntary_context_switches_inside_timed_calls": int(nivcsw), "steal_fraction": round(steal, 4),
               "loadavg_1min": load, "cpus": os.cpu_count()}
        why = []
        # a descheduled thread loses a scheduler slice (milliseconds) per switch; the timed calls last ~0.1-1 ms
        if nivcsw > 0 and nivcsw * 1e-3 > 0.05 * timed_seconds:
            why.append(f"host:timing thread descheduled {nivcsw}x")
        if steal > 0.02:
            why.append(f"host:steal {steal:.3f}")
        out["disturbed"] = why
        return out


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port)
# ------------------------------------------------------------------------------------------------

def cpu_check(offsets, lits, d, t, start, n_clauses, threads):
    from oracle_lib import check_db
    off = offsets[: n_clauses + 1]
    t0 = time.perf_counter()
    hits = check_db(off, lits[: off[-1]], d, t, start, use_filter=1, nthreads=threads, cap=1 << 20)
    return time.perf_counter() - t0, hits


def cpu_baseline(a, offsets, lits, d, t, start, seconds):
    cores = os.cpu_count() or 1
    probe = min(a.clauses, 200_000)
    dt, _ = cpu_check(offsets, lits, d, t, start, probe, cores)
    n = int(min(a.clauses, max(probe, probe * seconds / max(dt, 1e-3))))
    dt, hits = cpu_check(offsets, lits, d, t, start, n, cores)
    L = int(offsets[n])
    A = a.solvers * a.slots
    return {"value": L * A / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {n} clauses ({L} literals) of the workload x {A} assignments, "
                      f"two-level filter, {cores} threads, {dt:.2f} s"}, hits, n


def run_reference_arm(a):
    """--impl reference: the reference algorithm on the host cores (C port in oracle/)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # same workload as the own arm at this N (weak scaling: N x --clauses clauses); the sample is a
    # prefix of the database, so only the first --clauses clauses are generated
    total_clauses = a.clauses * (a.gpus if a.scaling == "weak" else 1)
    sig, offsets, lits = make_inputs(a)
    d, t, start = batch_words(a, sig)
    cores = os.cpu_count() or 1
    A = a.solvers * a.slots
    probe = min(a.clauses, 200_000)
    dt, _ = cpu_check(offsets, lits, d, t, start, probe, cores)
    budget = 150.0 / max(1, a.steps + a.warmup)  # whole run within a few minutes
    n = int(min(a.clauses, max(50_000, probe * min(budget, 20.0) / max(dt, 1e-3))))
    L = int(offsets[n])
    for _ in range(a.warmup):
        cpu_check(offsets, lits, d, t, start, n, cores)
    times = []
    for _ in range(a.steps):
        dt, _ = cpu_check(offsets, lits, d, t, start, n, cores)
        times.append(dt)
    total = sum(times)
    value = L * A * a.steps / total
    sample = (f"each step = first {n} of {total_clauses} clauses ({L} literals) x {A} assignments, "
              f"two-level filter, {cores} threads")
    a.clauses = total_clauses
    cfg = workload_config(a)
    if a.gpus > 1:
        cfg["clauses_per_gpu"] = total_clauses // a.gpus
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps, "higher_is_better": True, "scaling": a.scaling,
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def import_latency():
    """BASELINE config 5's shape (64 solver threads, clauses up to 200 literals) through the C ABI from
    C++ threads (tests/latency/latency_harness.cc): add a falsified clause -> the target solver pops it"""
    exe = os.path.join(ROOT, "tests", "latency", "latency_harness")
    if not os.path.exists(exe):
        return {"unavailable": "tests/latency/latency_harness not built"}
    out = {}
    # quiet: the solvers' assignments satisfy nearly every clause (a few background hits per run);
    # saturated: ~1000 background hits per solver and run, re-reported run after run
    for name, agree in (("quiet", "999"), ("saturated", "985")):
        try:
            r = subprocess.run([exe, "64", "200000", "1000000", "300", "-1", agree], capture_output=True, text=True, timeout=120)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
            out[name] = json.loads(line)
        except Exception as e:  # the headline numbers do not depend on it
            out[name] = {"unavailable": f"{type(e).__name__}: {e}"}
    return out


def workload_config(a):
    return {"workload": f"synthetic {a.clauses} clauses (Luby-like lengths 2-{a.max_len}) x {a.solvers * a.slots} "
                        f"assignments ({a.solvers} solvers x {a.slots} slots) over {a.vars} vars",
            "clauses": a.clauses, "vars": a.vars, "solvers": a.solvers, "slots": a.slots,
            "values": f"planted assignment, literal true/false/undef = {a.p_agree * (1 - a.p_undef):.3f}/"
                      f"{(1 - a.p_agree) * (1 - a.p_undef):.4f}/{a.p_undef}",
            "churn_per_assignment": a.churn,
            "l2_policy": "inputs larger than L2 (clause arenas + assignment tables > 126 MB)",
            "parallelism": ((f"clause tiles sharded x{a.gpus} ({a.scaling} scaling: {a.clauses} clauses in total), " +
                             ("one process, GPUSHARE_DEVICES: every device reads the batch from page-locked host memory and "
                              "writes its results there" if getattr(a, "devices", 1) > 1 else
                              f"assignments broadcast, exchange = {a.exchange}")) if a.gpus > 1 else "single GPU")}


# ------------------------------------------------------------------------------------------------
# reference GPU library arm (context only)
# ------------------------------------------------------------------------------------------------

def reference_gpu_numbers(a, sig, offsets, lits, steps, warmup):
    """the reference's own GPU library (oracle/_ref: gpuShareLib recompiled for sm_100a, unmodified) on
    the same inputs through the same call sequence: the kernel to beat"""
    import ref_lib
    if not ref_lib.available():
        return {"unavailable": "oracle/_ref/libgpushare_ref.so not built"}
    if a.solvers > 32:
        return {"unavailable": "the reference never checks solvers >= 32"}
    streams = make_streams(a, sig)
    sh = ref_lib.RefSharer(report=4000)
    sh.setVarCount(a.vars)
    sh.setCpuSolverCount(a.solvers)
    sh.addClausesBulk(offsets, lits)
    sh.gpuRun()
    L, A = int(offsets[-1]), a.solvers * a.slots
    pool = ThreadPoolExecutor(max_workers=a.solvers)
    times, kern = [], []
    for it in range(warmup + steps):
        push_batch(sh, streams, a.slots, pool)
        k0 = sh.getGlobalStat(9)  # timeSpentTestingClauses (us, its own CUDA events)
        t0 = time.perf_counter()
        sh.gpuRun(); sh.gpuRun()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
            kern.append(sh.getGlobalStat(9) - k0)
        for s in range(a.solvers):
            while sh.popReportedClause(s) is not None:
                pass
    sh.close()
    e2e = L * A * len(times) / sum(times)
    return {"e2e": {"value": e2e, "unit": UNIT, "ms_per_step": 1e3 * sum(times) / len(times)},
            "dFindClauses_us_per_step": float(np.mean(kern)),
            "value": L * A / (np.mean(kern) * 1e-6) if np.mean(kern) > 0 else None, "unit": UNIT,
            "steps": steps, "warmup": warmup,
            "note": "reference gpuShareLib recompiled for sm_100a (oracle/_ref), its own default grid (2 x SMs x 512), "
                    "same inputs, same gss_gpu_run x 2 call pattern; value = its own CUDA-event time of dFindClauses"}


def run_reference_gpu(a):
    sig, offsets, lits = make_inputs(a)
    r = reference_gpu_numbers(a, sig, offsets, lits, a.steps, a.warmup)
    r.update({"impl": "reference-gpu", "metric": METRIC, "config": workload_config(a)})
    if "e2e" in r:
        r["ms_per_step"] = r["e2e"]["ms_per_step"]
    print(json.dumps(r))


# ------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------

def run_b200_sharded(a):
    """N > 1: one process per GPU (torchrun).  The clause database is sharded by tiles: rank r
    checks its contiguous 1/N of the tiles of every length (every rank keeps the whole arena so that rank 0 can resolve
    any hit on its device).  Rank 0 is the front-end: it owns the solver streams, the assignment
    slot machines and the hand-over queues.

    --scaling weak (default): the database grows with N (N x --clauses clauses, i.e. --clauses per
    GPU) and every batch of 1024 assignments is checked against all of it -- what 8 x 180 GB are
    for.  --scaling strong: the --clauses database is split N ways (one batch is then ~100/N us of
    work per GPU: latency-bound by construction).

    --exchange peer (default): the batch and the hits travel over peer memory (CUDA IPC windows,
    peer loads / stores over NVLink, stream memory operations; csrc/peer.cu) -- no collective on the
    data path, NCCL only bootstraps.  --exchange nccl: one NCCL broadcast + one all-gather per batch
    (mgpu.ShardedRunner), kept for comparison."""
    import copy
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ["GPUSHARE_DEVICE"] = str(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    a.gpus = world
    # the line the driver reads is the --scaling one (default weak: per-GPU work fixed); the other
    # mode rides along as a sub-object measured in the same job, so SCALE carries both curves
    main = sharded_pass(copy.copy(a), a.scaling, a.steps, a.warmup, dist, torch, rank, world, local, device)
    other = "strong" if a.scaling == "weak" else "weak"
    if not a.no_other_scaling:
        sub = sharded_pass(copy.copy(a), other, max(3, a.steps // 2), a.warmup, dist, torch, rank, world, local, device)
        if rank == 0:
            main[other + "_scaling"] = {k: sub[k] for k in ("value", "unit", "ms_per_step", "scaling", "steps", "e2e", "hits_per_step",
                                                            "phases_us_per_step", "parity_sample", "literals", "e2e_host_us_per_step_rank0") if k in sub}
            main[other + "_scaling"]["clauses"] = sub["config"]["clauses"]
    if rank == 0:
        print(json.dumps(main))
    dist.barrier()
    dist.destroy_process_group()


def sharded_pass(a, scaling, steps, warmup, dist, torch, rank, world, local, device):
    from gpusharesat_b200 import GpuClauseSharer, GpuClauseSharerOptions, mgpu
    a.scaling, a.steps, a.warmup = scaling, steps, warmup
    per_gpu = a.clauses
    if a.scaling == "weak":
        a.clauses = per_gpu * world
    sig, offsets, lits = make_inputs(a)
    L_total, A = int(offsets[-1]), a.solvers * a.slots
    sh = GpuClauseSharer(GpuClauseSharerOptions(minGpuLatencyMicros=0, verbosity=0))
    sh.setShard(rank, world)
    sh.setVarCount(a.vars)
    sh.setCpuSolverCount(a.solvers)
    sh.addClausesBulk(offsets, lits)
    # rank 0 keeps a prefix of the clause stream for the parity sample (ids = positions in the stream; the
    # arenas are sorted by first literal, so any id range is spread over every rank's share of the tiles)
    n_sample = min(a.clauses, a.parity_clauses)
    s_off = offsets[: n_sample + 1].copy() if rank == 0 else None
    s_lits = lits[: int(offsets[n_sample])].copy() if rank == 0 else None
    del offsets, lits
    streams = make_streams(a, sig) if rank == 0 else None
    pool = ThreadPoolExecutor(max_workers=a.solvers) if rank == 0 else None

    def drain():
        if rank == 0:
            for s in range(a.solvers):
                while sh.popReportedClause(s) is not None:
                    pass

    # the first batch rebuilds the tables: it lists every variable of every solver
    payload_cap = a.solvers * a.vars * 12 + (4 << 20)
    if a.exchange == "peer":
        # the parity batch (first warm-up step) needs every rank's masks on rank 0; the timed steps run the default
        runner = mgpu.PeerRunner(sh, dist, rank, world, payload_cap=payload_cap, slot_hits=a.slot_hits, records=True)
    else:
        runner = mgpu.ShardedRunner(sh, dist, rank, world, device, payload_cap=payload_cap)

    def step():
        if rank == 0:
            push_batch(sh, streams, a.slots, pool)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nh = runner.step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ph = sh.debugLastRunTimes()
        h2d, d2h = sh.debugLastRunBytes()
        dev_us = runner.device_us()
        drain()  # (gss_debug_last_hits stays valid: it reads the run's records, not the solver queues)
        return dt, ph, (nh or 0) if rank == 0 else 0, h2d, d2h, dev_us

    first_hits = None
    for w in range(a.warmup):
        step_out = step()
        if w == 0 and rank == 0:
            first_hits = sh.debugLastHits().copy()  # the union of every rank's hits of the first batch
        if w == 0 and a.exchange == "peer":
            runner.set_records(False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = sh.debugKernelLaunches()
    hp0 = sh.debugHostPhases()
    wall = dev = 0.0
    hits = upd = back = 0
    ex, tk, ck = [], [], []
    for _ in range(a.steps):
        dt, ph, nh, h2d, d2h, dev_us = step()
        wall += dt
        dev += dev_us * 1e-6
        ex.append(dev_us - ph[1] - ph[2]); tk.append(ph[1]); ck.append(ph[2])
        hits += nh
        upd += h2d
        back += d2h
    launches = sh.debugKernelLaunches() - l0
    hp = [(b - a0) / a.steps for a0, b in zip(hp0, sh.debugHostPhases())]
    clocks = sampler.stop() if rank == 0 else None
    tt = torch.tensor([dev, wall], dtype=torch.float64, device=device)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ll = torch.tensor([launches], dtype=torch.int64, device=device)
    dist.all_reduce(ll)
    mine = torch.tensor([float(np.mean(tk)), float(np.mean(ck)), float(np.mean(ex))], dtype=torch.float64, device=device)
    allph = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allph, mine)
    dev, wall = float(tt[0]), float(tt[1])
    out = None
    if rank == 0:
        parity = None
        if first_hits is not None and not a.no_cpu:
            from oracle_lib import check_db
            d, t, start = batch_words(a, sig)
            cpu_hits = check_db(s_off, s_lits, d, t, start, use_filter=1, nthreads=os.cpu_count() or 1, cap=1 << 22)
            g = first_hits[first_hits["clause_id"] < n_sample]
            parity = {"clauses": int(n_sample), "gpu_hits": int(len(g)), "cpu_hits": int(len(cpu_hits)),
                      "identical": bool(np.array_equal(g, cpu_hits)),
                      "note": "first batch: union of all ranks' hits restricted to the first clauses of the stream (spread over "
                              "every rank's tiles by the first-literal sort) == the CPU oracle on those clauses"}
        cfg = workload_config(a)
        cfg["clauses_per_gpu"] = a.clauses // world
        if a.exchange == "peer":
            region = ("every rank, on its own stream (CUDA events): batch resident in its HBM (rank 0: after the H2D; workers: "
                      "after rank 0's push over NVLink has arrived) -> table kernels -> check kernels on the rank's share of "
                      "the tiles -> per-solver sort and emission of the rank's finished results (ids, literal stream) into "
                      "host memory over the rank's own PCIe link, plus the collapse; max over ranks")
            gap = "exchange_and_wait_for_slowest_rank"
        else:
            region = "NCCL broadcast of the batch, table + check kernels on every shard, NCCL all-gather of the hits; max over ranks"
            gap = "nccl_broadcast_gather_and_sync_gaps"
        out = ({
            "metric": METRIC, "value": L_total * A * a.steps / dev, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * dev / a.steps, "higher_is_better": True, "scaling": a.scaling,
            "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": cfg,
            "e2e": {"value": L_total * A * a.steps / wall, "unit": UNIT, "ms_per_step": 1e3 * wall / a.steps,
                    "h2d_bytes_per_step": int(upd / a.steps),
                    "d2h_bytes_per_step": int(back / a.steps),
                    "timed_region": "rank 0 collect + H2D, push to the workers, table + check + emit kernels on every rank (each "
                                    "rank writes its finished per-solver results into shared host memory), rank 0 stitches "
                                    "the per-rank slices into the solvers' batches (zero-copy); wall clock, max over ranks"},
            "gpu_launches": int(ll[0]), "clocks": clocks, "hits_per_step": hits / a.steps,
            "phases_us_per_step": {gap: float(np.mean(ex)), "table_kernels": float(np.mean(tk)),
                                   "check_kernels": float(np.mean(ck)),
                                   "per_rank_table_check_other": [[round(float(x), 1) for x in t.tolist()] for t in allph]},
            "literals": L_total, "assignments": A, "parity_sample": parity,
            "e2e_host_us_per_step_rank0": {"enqueue": hp[1], "of_which_collect_deltas": hp[3], "finish": hp[0],
                                           "of_which_wait_for_own_gpu": hp[4], "of_which_wait_for_workers": hp[5],
                                           "of_which_hand_over": hp[2]},
            "note": "N > 1: value = " + region + "; e2e adds rank 0's collect, the payload H2D, the wait for the slowest rank's "
                    "publication and the host hand-over",
        })
    dist.barrier()
    del runner
    sh.close()
    dist.barrier()
    return out


def streamed_db_numbers(a, sig, offsets, lits, chunk):
    """SURVEY 8(f1) / BASELINE configs[3] shape on this workload's database: the SAME clauses arrive `chunk` at a
    time with a run between the chunks (2 fresh assignments per solver), the way solver threads feed the library --
    never as one bulk load.  Reports the ingest rate including the runs, the distribution of gpuRun() wall times
    while the arenas grow (in place: vmem.cc), the device-side re-sorts of the streamed tails (reduce.cu), what the
    level-1 kernel costs on the streamed database against the bulk-loaded one, and reduceDb."""
    from gpusharesat_b200 import GpuClauseSharer, GpuClauseSharerOptions, GlobalStats
    n = len(offsets) - 1
    sh = GpuClauseSharer(GpuClauseSharerOptions(minGpuLatencyMicros=0, verbosity=0))
    sh.setVarCount(a.vars)
    sh.setCpuSolverCount(a.solvers)
    streams = make_streams(a, sig)
    pool = ThreadPoolExecutor(max_workers=a.solvers)

    def drain():
        for s in range(a.solvers):
            while sh.popReportedClause(s) is not None:
                pass

    run_ms, t_all, t_add, resort_at = [], 0.0, 0.0, []
    sh.debugHostPhases()  # (GSS_HOST_PROF: start of this pass)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        off = offsets[lo:hi + 1] - offsets[lo]
        part = lits[offsets[lo]:offsets[hi]]
        t0 = time.perf_counter()
        sh.addClausesBulk(off, part)
        t_add += time.perf_counter() - t0
        push_batch(sh, streams, 2, pool)
        t1 = time.perf_counter()
        sh.gpuRun()
        t2 = time.perf_counter()
        drain()
        run_ms.append((t2 - t1) * 1e3)
        resort_at.append(sh.debugDbOrder()[1])
        t_all += t2 - t0
    sh.gpuRun()
    drain()
    unsorted, resorts = sh.debugDbOrder()
    push_batch(sh, streams, a.slots, pool)
    sh.gpuRun()
    t_filter = sh.debugTimeCheck(a.prod_iters, filter_only=True)
    t_prod = sh.debugTimeCheck(a.prod_iters, dense=False)
    sh.gpuRun()
    drain()
    before = sh.getGlobalStat(GlobalStats.gpuClauses)
    sh.debugHostPhases()  # (GSS_HOST_PROF: the streaming part ends here)
    t0 = time.perf_counter()
    sh.reduceDb()
    t_reduce = time.perf_counter() - t0
    after = sh.getGlobalStat(GlobalStats.gpuClauses)
    push_batch(sh, streams, a.slots, pool)
    t0 = time.perf_counter()
    sh.gpuRun(); sh.gpuRun()
    t_after = time.perf_counter() - t0
    drain()
    oom = bool(sh.hasRunOutOfGpuMemoryOnce())
    sh.close()
    sh_phases = None
    # the first run builds the assignment tables from scratch (every variable of every solver): reported on its own
    r = np.array(run_ms[1:]) if len(run_ms) > 1 else np.array(run_ms)
    with_resort = [run_ms[i] for i in range(1, len(run_ms)) if resort_at[i] != resort_at[i - 1]]
    return {"workload": f"the same {n} clauses arriving {chunk} per run, {a.solvers} solvers x 2 assignments per run",
            "ingest_clauses_per_s_incl_runs": n / t_all,
            "add_clauses_bulk_ms_per_chunk": 1e3 * t_add / max(1, len(run_ms)),
            "first_run_ms_table_build": run_ms[0],
            "gpu_run_ms_while_growing": {"p50": float(np.percentile(r, 50)), "p90": float(np.percentile(r, 90)),
                                         "p99": float(np.percentile(r, 99)), "max": float(r.max()), "runs": len(r),
                                         "runs_with_a_resort_ms": [round(x, 2) for x in with_resort],
                                         "all_ms": [round(x, 2) for x in run_ms]},
            "device_resorts": int(resorts), "unsorted_clauses_at_the_end": int(unsorted),
            "k_filter_us": t_filter, "check_kernels_us": t_prod,
            "reduce_db_ms": t_reduce * 1e3, "clauses_before_after_reduce": [int(before), int(after)],
            "first_batch_after_reduce_ms": t_after * 1e3, "out_of_memory": oom}


def run_b200(a):
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return run_b200_sharded(a)
    from gpusharesat_b200 import GpuClauseSharer, GpuClauseSharerOptions
    rank, world, local, dist = 0, 1, 0, None
    os.environ.setdefault("GPUSHARE_DEVICE", "0")
    if a.devices > 1:  # one process, N devices behind the one sharer
        os.environ["GPUSHARE_DEVICES"] = str(a.devices)
        a.gpus = a.devices
        if a.scaling == "weak":
            a.clauses *= a.devices
        a.no_dense = a.no_ref_gpu = True

    sig, offsets, lits = make_inputs(a)
    L_total, A = int(offsets[-1]), a.solvers * a.slots
    my_off, my_lits = offsets, lits

    sh = GpuClauseSharer(GpuClauseSharerOptions(minGpuLatencyMicros=0, verbosity=0))
    sh.setVarCount(a.vars)
    sh.setCpuSolverCount(a.solvers)
    sh.addClausesBulk(my_off, my_lits)
    streams = make_streams(a, sig)  # identical on every rank: the broadcast payload is replayed locally
    pool = ThreadPoolExecutor(max_workers=a.solvers)

    def barrier():
        if dist is not None:
            dist.barrier()

    def one_step():
        """returns (wall seconds of execute(), device us of this step's run, hits)"""
        push_batch(sh, streams, a.slots, pool)
        barrier()
        t0 = time.perf_counter()
        sh.gpuRun()
        sh.gpuRun()
        dt = time.perf_counter() - t0
        ph = sh.debugLastRunTimes()
        return dt, ph, len(sh.debugLastHits())

    def drain():
        for s in range(a.solvers):
            while sh.popReportedClause(s) is not None:
                pass

    first_hits = None
    for w in range(a.warmup):
        dt, ph, nh = one_step()
        if w == 0:
            first_hits = sh.debugLastHits().copy()
        drain()

    def timed_region():
        """W warm-up steps are done; times exactly a.steps steps.  Also watches the HOST while it times: a thread
        that was descheduled, or vCPU time stolen by the hypervisor, makes the wall-clock (e2e) number a
        measurement of the neighbours, not of the path."""
        sampler = ClockSampler(local)
        sampler.start()
        launches0 = sh.debugKernelLaunches()
        hp0 = sh.debugHostPhases()
        wall, dev_tables, dev_check, dev_total, hits, h2d, d2h = [], [], [], [], [], [], []
        collapse_us = []
        host0 = HostWatch()
        barrier()
        for it in range(a.steps):
            push_batch(sh, streams, a.slots, pool)
            barrier()
            host0.enter()
            t0 = time.perf_counter()
            sh.gpuRun()                       # starts this batch's run (and gathers the empty run before it)
            prev = sh.debugLastRunTimes()     # phases of the empty run: holds the deferred collapse of the previous batch
            b1 = sh.debugLastRunBytes()
            sh.gpuRun()                       # gathers it: hits are on the host, handed to the solver queues
            dt = time.perf_counter() - t0
            host0.leave()
            ph = sh.debugLastRunTimes()
            wall.append(dt)
            collapse_us.append(prev[1] if prev else 0.0)
            dev_tables.append(ph[1]); dev_check.append(ph[2]); dev_total.append(ph[3])
            hits.append(len(sh.debugLastHits()))
            h2d.append(b1[0]); d2h.append(sh.debugLastRunBytes()[1])
            drain()
        host = host0.stop(sum(wall))
        launches = sh.debugKernelLaunches() - launches0
        hp = [(b - a0) / a.steps for a0, b in zip(hp0, sh.debugHostPhases())]
        clocks = sampler.stop()
        return dict(wall=wall, dev_tables=dev_tables, dev_check=dev_check, dev_total=dev_total, hits=hits, h2d=h2d, d2h=d2h,
                    collapse_us=collapse_us, launches=launches, hp=hp, clocks=clocks, host=host)

    # Re-measure ONCE when the region was disturbed -- the GPU throttled (contract) or the host thread lost its
    # core (involuntary context switches of the timing thread / stolen vCPU time); both attempts are reported.
    attempts = []
    for attempt in range(2):
        r = timed_region()
        bad = [x for x in r["clocks"].get("reasons", []) if x != "sw_power_cap"]
        r["rejected"] = (["gpu:" + x for x in bad] + r["host"]["disturbed"]) if attempt == 0 else []
        attempts.append(r)
        if not r["rejected"]:
            break
    r = attempts[-1]
    wall, dev_tables, dev_check, dev_total, hits, h2d, d2h = (r[k] for k in ("wall", "dev_tables", "dev_check", "dev_total", "hits", "h2d", "d2h"))
    collapse_us, launches, hp, clocks = r["collapse_us"], r["launches"], r["hp"], r["clocks"]
    # the collapse of batch k runs at the start of the following run: charge it to batch k
    sh.gpuRun()
    tail = sh.debugLastRunTimes()
    collapse = collapse_us[1:] + [tail[1] if tail else 0.0]
    dev_step_us = [t + c + k for t, c, k in zip(dev_tables, dev_check, collapse)]

    dev_s = sum(dev_step_us) * 1e-6
    wall_s = sum(wall)
    if dist is not None:
        import torch
        tt = torch.tensor([dev_s, wall_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_s, wall_s = float(tt[0]), float(tt[1])
        th = torch.tensor([sum(hits)], dtype=torch.int64, device="cuda")
        dist.all_reduce(th)
        total_hits = int(th[0])
    else:
        total_hits = sum(hits)

    out = {
        "metric": METRIC, "value": L_total * A * a.steps / dev_s, "unit": UNIT, "n_gpus": max(world, a.devices), "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * dev_s / a.steps, "higher_is_better": True, "scaling": a.scaling,
        "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": workload_config(a),
        "e2e": {"value": L_total * A * a.steps / wall_s, "unit": UNIT, "ms_per_step": 1e3 * wall_s / a.steps,
                "ms_every_step": [round(1e3 * w, 3) for w in wall],
                "h2d_bytes_per_step": int(np.mean(h2d)), "d2h_bytes_per_step": int(np.mean(d2h)),
                "timed_region": "gss_gpu_run() x 2 per batch: collect (buffer swap), header H2D, k_apply_direct reading the deltas "
                                "from page-locked host memory, check kernels, k_emit writing ids + literals into page-locked "
                                "host memory, zero-copy hand-over to the solver queues"},
        "gpu_launches": int(launches), "clocks": clocks,
        "host_during_timed_region": r["host"],
        "remeasured": ([{"rejected_because": x["rejected"], "e2e_ms_per_step": 1e3 * sum(x["wall"]) / a.steps,
                         "device_ms_per_step": 1e-3 * float(np.mean(x["dev_tables"]) + np.mean(x["dev_check"]))}
                        for x in attempts[:-1]] or None),
        "hits_per_step": total_hits / a.steps,
        "phases_us_per_step": {"table_kernels": float(np.mean(dev_tables) + np.mean(collapse)),
                               "check_kernels": float(np.mean(dev_check)), "h2d_to_d2h_total": float(np.mean(dev_total))},
        # everything the device does for one batch, first H2D byte to last D2H byte (CUDA events)
        "device_step_complete": {"us": float(np.mean(dev_total) + np.mean(collapse)),
                                 "value": L_total * A / ((float(np.mean(dev_total)) + float(np.mean(collapse))) * 1e-6), "unit": UNIT},
        "literals": L_total, "assignments": A,
        "e2e_host_us_per_step": {"finish_previous_run": hp[0], "of_which_wait_for_gpu": hp[4],
                                 "of_which_sort_resolve_d2h_of_hits": hp[5], "start_next_run": hp[1],
                                 "of_which_collect_deltas": hp[3], "hand_over_to_solver_queues": hp[2]},
        "host_us_total": {"fill_assigs": sh.getGlobalStat(10), "fill_reported": sh.getGlobalStat(11),
                          "gpu_runs": sh.getGlobalStat(3)},
    }

    if a.devices > 1:
        out["phases_us_per_step"]["note"] = "device times = the slowest device's (max over the devices of every phase)"
        if not a.no_cpu and first_hits is not None:
            from oracle_lib import check_db
            d, t, start = batch_words(a, sig)
            n_sample = min(a.clauses, a.parity_clauses)
            off = offsets[: n_sample + 1]
            cpu_hits = check_db(off, lits[: off[-1]], d, t, start, use_filter=1, nthreads=os.cpu_count() or 1, cap=1 << 22)
            g = first_hits[first_hits["clause_id"] < n_sample]
            out["parity_sample"] = {"clauses": int(n_sample), "gpu_hits": int(len(g)), "cpu_hits": int(len(cpu_hits)),
                                    "identical": bool(np.array_equal(g, cpu_hits))}
        print(json.dumps(out))
        return
    if rank == 0 and world == 1:
        # kernel-only timings on the last batch (tables still resident)
        push_batch(sh, streams, a.slots, pool)
        sh.gpuRun()
        t_prod = sh.debugTimeCheck(a.prod_iters, dense=False)
        t_filter = sh.debugTimeCheck(a.prod_iters, filter_only=True)
        out["kernel_us"] = {"k_filter": t_filter, "k_exact": sh.debugTimeCheck(a.prod_iters, mode=3),
                            "k_apply_updates": sh.debugTimeCheck(a.prod_iters, mode=4),
                            "k_collapse": sh.debugTimeCheck(a.prod_iters, mode=5),
                            "k_emit": sh.debugTimeCheck(a.prod_iters, mode=6),
                            "note": "each kernel re-launched back to back on the last batch (tables resident), CUDA events"}
        if a.filter_sweep:
            names = sh.debugFilterVariants()
            sweep = []
            for v, name in enumerate(names):
                sh.debugSetFilterVariant(v)
                sweep.append({"variant": v, "name": name, "us": sh.debugTimeCheck(a.prod_iters, filter_only=True)})
            sh.debugSetFilterVariant(int(os.environ.get("GSS_FILTER_VARIANT", "-1")))
            out["filter_variants"] = sweep
            sweep = []
            for v, name in enumerate(sh.debugExactVariants()):
                sh.debugSetExactVariant(v)
                sweep.append({"variant": v, "name": name, "us": sh.debugTimeCheck(a.prod_iters, mode=3)})
            sh.debugSetExactVariant(int(os.environ.get("GSS_EXACT_VARIANT", "-1")))
            out["exact_variants"] = sweep
        lop3 = sh.debugLop3Peak()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        # DRAM traffic per launch: ncu cannot run inside a timed bench, so the numbers come from the
        # capture profiles/capture_traffic.py makes of THIS command (regenerated whenever a kernel changes)
        traffic, traffic_src = {}, None
        for name in ("r02_traffic.json", "r01_traffic.json"):
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", name)))
                traffic, traffic_src = tj["bytes_per_launch"], f"profiles/{name}: " + tj.get("source", "")
                break
            except Exception:
                pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        W = (A + 31) // 32
        out["kernel_production"] = {"us_per_sweep": t_prod, "checks_per_s": L_total * A / (t_prod * 1e-6),
                                    "kernels": "k_filter + k_exact", "k_filter_us": t_filter}
        # ---- k_filter (level 1), the dominant kernel of the production step, against measured HBM ----
        # nominal bytes = 4 B x every literal + the level-1 table once (16 B per variable); the kernel
        # legitimately moves LESS through DRAM (`traffic`, ncu): whole tiles die early and their remaining
        # rows are skipped -- both fractions are reported.
        fb = 4.0 * L_total + 16.0 * a.vars
        step_us = 1e6 * dev_s / a.steps
        moved = traffic.get("k_filter")
        out["roofline_k_filter"] = {
            "kernel": "k_filter", "bound": "hbm", "peak": hbm_peak, "unit": "GB/s", "us_per_launch": t_filter,
            "nominal_bytes": fb, "nominal_gbs": fb / (t_filter * 1e-6) / 1e9, "frac_nominal": fb / (t_filter * 1e-6) / 1e9 / hbm_peak,
            "moved_bytes": moved, "frac_moved": (moved / (t_filter * 1e-6) / 1e9 / hbm_peak) if moved else None,
            "share_of_step": t_filter / step_us, "peak_source": peak_src, "traffic_source": traffic_src,
            "whole_step": {"algorithmic_bytes": fb + 2 * float(np.mean(h2d)) + 16.0 * float(np.mean(hits)), "us": step_us,
                           "frac": (fb + 2 * float(np.mean(h2d)) + 16.0 * float(np.mean(hits))) / (step_us * 1e-6) / 1e9 / hbm_peak}}
        if not a.no_dense:
            t_dense = sh.debugTimeCheck(a.dense_iters, dense=True)
            t_slice = sh.debugTimeCheck(a.dense_iters, mode=7)
            sh.gpuRun()
            n_dense = len(sh.debugLastHits())
            drain()
            H = n_dense
            bytes_alg = 4.0 * L_total + 8.0 * a.vars * W + 16.0 * H
            lop3_alg = 3.0 * L_total * W
            t_hbm, t_int = bytes_alg / (hbm_peak * 1e9), lop3_alg / lop3
            bound_int = t_int >= t_hbm
            # BASELINE.json's own definition: roofline = slower of HBM bytes and LOP3 issue with EVERY
            # (literal, 32-slot word) pair evaluated -- a bench-only mode (no filter, no early exit)
            out["roofline"] = {
                "kernel": "k_check_dense", "mode": "dense (no filter, no early exit): the contract's roofline (SURVEY 8d)",
                "bound": "int_lop3" if bound_int else "hbm",
                "achieved": (lop3_alg / (t_dense * 1e-6)) / 1e12 if bound_int else bytes_alg / (t_dense * 1e-6) / 1e9,
                "peak": lop3 / 1e12 if bound_int else hbm_peak,
                "unit": "TLOP3/s" if bound_int else "GB/s",
                "frac": max(t_hbm, t_int) / (t_dense * 1e-6),
                # (sliced variant: one launch per slice of 8 solvers -> W/8... = ceil(solvers / 8) launches per sweep)
                "traffic": (traffic.get("k_check_dense_sliced") * ((min(a.solvers, 32) + 7) // 8)
                            if traffic.get("k_check_dense_sliced") and not os.environ.get("GSS_DENSE_FLAT") and a.solvers > 8
                            else traffic.get("k_check_dense")),
                "traffic_source": traffic_src,
                "us_per_sweep": t_dense, "t_roof_us": max(t_hbm, t_int) * 1e6,
                "t_hbm_us": t_hbm * 1e6, "t_int_us": t_int * 1e6,
                "algorithmic_bytes": bytes_alg, "algorithmic_lop3": lop3_alg,
                "hbm_view": {"achieved_gbs": bytes_alg / (t_dense * 1e-6) / 1e9, "peak_gbs": hbm_peak, "peak_source": peak_src},
                "lop3_peak_source": "measured in this run (register-only LOP3 micro-benchmark, gss_debug_lop3_peak)",
                "gather_view": {"table_bytes_gathered": 8.0 * L_total * W,
                                "gathered_gbs": 8.0 * L_total * W / (t_dense * 1e-6) / 1e9,
                                "l2_gather_ceiling_gbs": L2_GATHER_GBS,
                                "t_l2_gather_us": 8.0 * L_total * W / (L2_GATHER_GBS * 1e9) * 1e6,
                                "frac_of_l2_gather_ceiling": (8.0 * L_total * W / (L2_GATHER_GBS * 1e9)) / (t_dense * 1e-6),
                                "note": "every (literal, word) pair gathers 8 B of table data that has no reuse "
                                        "structure (random variables): 8*L*W bytes have to cross L2 -> SM whatever the "
                                        "layout.  The ceiling is the measured rate of 64-byte row gathers from an "
                                        "L2-resident 64 MB table on this GPU (profiles/r02_probes.jsonl, probe 'gather', "
                                        "row64_u8); it, not LOP3 issue or HBM, bounds dense mode"},
                "kernel_variant": ("k_check_dense_sliced: table cut into 8-solver slices of 64 MB that stay in L2, one sweep "
                                   "of the clauses per slice" if not os.environ.get("GSS_DENSE_FLAT") and a.solvers > 8
                                   else "k_check_dense: 256-byte rows of the whole table (mostly from HBM)"),
                "slice_build_us_per_batch": t_slice,
                "dense_hits": n_dense,
                "production_speedup_over_dense": t_dense / t_prod,
                # the production kernels do the same NOMINAL checks in t_prod: relative to the dense roofline
                # time they are > 1 because the two-level filter never evaluates provably dead pairs
                "production_us_per_sweep": t_prod, "t_roof_over_production": max(t_hbm, t_int) * 1e6 / t_prod,
            }
        else:
            sh.gpuRun()
            drain()
        sh.close()  # free the device: the latency harness and the reference library bring their own databases
        if not a.no_streamed:
            out["streamed_db"] = streamed_db_numbers(a, sig, offsets, lits, a.stream_chunk)
            out["streamed_db"]["k_filter_us_bulk_loaded"] = t_filter
        if not a.no_latency:
            out["import_latency"] = import_latency()
        if not a.no_ref_gpu:
            out["reference_gpu"] = reference_gpu_numbers(a, sig, offsets, lits, a.ref_gpu_steps, 2)
            if "e2e" in out["reference_gpu"]:
                out["reference_gpu"]["ours_over_reference"] = {
                    "e2e": out["e2e"]["value"] / out["reference_gpu"]["e2e"]["value"],
                    "check_kernels": out["reference_gpu"]["dFindClauses_us_per_step"] / float(np.mean(dev_check))}
        if not a.no_cpu:
            d, t, start = batch_words(a, sig)
            cb, cpu_hits, n_sample = cpu_baseline(a, offsets, lits, d, t, start, a.cpu_seconds)
            out["cpu_baseline"] = cb
            # full-size parity on the sample: the first batch's GPU hits restricted to the sampled clauses
            if first_hits is not None:
                g = first_hits[first_hits["clause_id"] < n_sample]
                out["parity_sample"] = {"clauses": int(n_sample), "gpu_hits": int(len(g)), "cpu_hits": int(len(cpu_hits)),
                                        "identical": bool(np.array_equal(g, cpu_hits))}
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    elif a.impl == "reference-gpu":
        run_reference_gpu(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
