"""Test-side binding of the CPU oracle (oracle/gss_oracle.c) plus a snapshot model of the
sharer API built on it.  TEST INFRASTRUCTURE ONLY -- the product never imports this."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
HIT_DTYPE = np.dtype([("clause_id", "<i8"), ("solver_id", "<i4"), ("mask", "<u4")])
_LIB = None


def build_oracle():
    so = os.path.join(ORACLE_DIR, "libgss_oracle.so")
    src = os.path.join(ORACLE_DIR, "gss_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "libgss_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def oracle():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build_oracle())
        L.gss_oracle_kat_run.restype = C.c_int64
        L.gss_oracle_kat_run.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64)]
        L.gss_oracle_kat_clauses.restype = C.c_int64
        L.gss_oracle_kat_clauses.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_void_p, C.c_void_p]
        L.gss_oracle_kat_assignment.restype = None
        L.gss_oracle_kat_assignment.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_void_p]
        L.gss_oracle_clause_fires.restype = C.c_int
        L.gss_oracle_clause_fires.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.gss_oracle_clause_mask32.restype = C.c_uint32
        L.gss_oracle_clause_mask32.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32]
        L.gss_oracle_check_db.restype = C.c_int64
        L.gss_oracle_check_db.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int]
        L.gss_oracle_check_db_bench.restype = C.c_int64
        L.gss_oracle_check_db_bench.argtypes = L.gss_oracle_check_db.argtypes[:10] + [C.c_int]
        _LIB = L
    return _LIB


def check_db(offsets, lits, def_, tru, start, use_filter=0, nthreads=1, cap=None, bench=False):
    """def_/tru: uint32 [nsolvers, nvars]; returns sorted hit array (clause index as clause_id)"""
    L = oracle()
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    lits = np.ascontiguousarray(lits, dtype=np.int32)
    def_ = np.ascontiguousarray(def_, dtype=np.uint32)
    tru = np.ascontiguousarray(tru, dtype=np.uint32)
    start = np.ascontiguousarray(start, dtype=np.uint32)
    nsolvers, nvars = def_.shape
    ncl = offsets.size - 1
    if cap is None:
        cap = 1 << 16
    while True:
        out = np.zeros(cap, dtype=HIT_DTYPE)
        if bench:  # bench.py's CPU arm: two-level filter, aggregates built by all threads
            n = L.gss_oracle_check_db_bench(offsets.ctypes.data, lits.ctypes.data, ncl, nsolvers, nvars, def_.ctypes.data,
                                            tru.ctypes.data, start.ctypes.data, out.ctypes.data, cap, nthreads)
        else:
            n = L.gss_oracle_check_db(offsets.ctypes.data, lits.ctypes.data, ncl, nsolvers, nvars, def_.ctypes.data,
                                      tru.ctypes.data, start.ctypes.data, out.ctypes.data, cap, use_filter, nthreads)
        if n <= cap:
            out = out[:n]
            break
        cap = int(n)
    return np.sort(out, order=["clause_id", "solver_id"])


def kat_clauses(nclauses, min_len, max_len, nvars, seed=0.4):
    L = oracle()
    offsets = np.zeros(nclauses + 1, dtype=np.int64)
    lits = np.zeros(nclauses * max_len, dtype=np.int32)
    s = C.c_double(seed)
    n = L.gss_oracle_kat_clauses(nclauses, min_len, max_len, nvars, C.byref(s), offsets.ctypes.data, lits.ctypes.data)
    return offsets, lits[:n].copy()


class KatAssignments:
    """perfTest.cu:70-83 assignment stream (seed 0.6 carried across iterations)"""

    def __init__(self, nvars, seed=0.6):
        self.nvars = nvars
        self.seed = C.c_double(seed)

    def next(self):
        vals = np.zeros(self.nvars, dtype=np.uint8)
        oracle().gss_oracle_kat_assignment(C.byref(self.seed), self.nvars, vals.ctypes.data)
        return vals


class SharerModel:
    """Snapshot model of the GpuClauseSharer API: a slot is the solver's whole partial assignment
    frozen at trySendAssignment time (SURVEY.md Appendix A rules 1-4); every run tests every clause
    of the database against every slot frozen since the previous run with the scalar oracle
    semantics.  Independent of the delta / collapse machinery it checks."""

    def __init__(self, nvars, nsolvers):
        self.nvars = nvars
        self.S = nsolvers
        self.vals = [np.full(nvars, 2, dtype=np.uint8) for _ in range(nsolvers)]
        self.to_unset = [[] for _ in range(nsolvers)]
        self.first = [0] * nsolvers
        self.cur = [0] * nsolvers
        self.snaps = [dict() for _ in range(nsolvers)]
        self.db = []       # (clause_id, lits)
        self.pending = []
        self.next_id = 0
        self.max_len = 100

    def _avail(self, s):
        return self.cur[s] != self.first[s] + 32

    def _flush(self, s):
        for l in self.to_unset[s]:
            self.vals[s][l >> 1] = 2
        self.to_unset[s] = []

    def trySetSolverValues(self, s, lits):
        if not self._avail(s):
            return False
        self._flush(s)
        for l in lits:
            self.vals[s][l >> 1] = 1 if (l & 1) else 0
        return True

    def unsetSolverValues(self, s, lits):
        if self._avail(s):
            self._flush(s)
            for l in lits:
                self.vals[s][l >> 1] = 2
        else:
            self.to_unset[s].extend(lits)

    def trySendAssignment(self, s):
        if not self._avail(s):
            return -1
        self.snaps[s][self.cur[s]] = self.vals[s].copy()
        self.cur[s] += 1
        return self.cur[s] - 1

    def addClause(self, lits):
        if len(lits) > self.max_len or len(lits) < 1:
            return -1
        self.pending.append((self.next_id, list(lits)))
        self.next_id += 1
        return self.next_id - 1

    def run(self):
        """one started GPU run; returns sorted hits or None when the reference would not start one"""
        self.db.extend(self.pending)
        self.pending = []
        if not self.db:
            return None
        def_ = np.zeros((self.S, self.nvars), dtype=np.uint32)
        tru = np.zeros((self.S, self.nvars), dtype=np.uint32)
        start = np.zeros(self.S, dtype=np.uint32)
        for s in range(self.S):
            for i in range(self.first[s], self.cur[s]):
                p = np.uint32(1 << (i % 32))
                v = self.snaps[s][i]
                def_[s][v != 2] |= p
                tru[s][v == 0] |= p
                start[s] |= p
            self.snaps[s] = {}
            self.first[s] = self.cur[s]
        offsets = np.zeros(len(self.db) + 1, dtype=np.int64)
        flat = []
        for k, (_, l) in enumerate(self.db):
            flat.extend(l)
            offsets[k + 1] = len(flat)
        hits = check_db(offsets, np.array(flat, dtype=np.int32), def_, tru, start)
        ids = np.array([cid for cid, _ in self.db], dtype=np.int64)
        hits["clause_id"] = ids[hits["clause_id"]]
        return np.sort(hits, order=["clause_id", "solver_id"])
