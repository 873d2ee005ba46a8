"""Test-side binding of oracle/_ref/libgpushare_ref.so: the UNMODIFIED reference gpuShareLib
compiled for sm_100a plus oracle/ref_harness.cu.  TEST INFRASTRUCTURE ONLY.  Same method names
as gpusharesat_b200.GpuClauseSharer so one test body drives both."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libgpushare_ref.so")
HIT_DTYPE = np.dtype([("clause_id", "<i8"), ("solver_id", "<i4"), ("mask", "<u4")])
_IP = C.POINTER(C.c_int)
_LIB = None


class ref_options(C.Structure):
    _fields_ = [("gpuBlockCountGuideline", C.c_int), ("gpuThreadsPerBlockGuideline", C.c_int),
                ("minGpuLatencyMicros", C.c_int), ("verbosity", C.c_int),
                ("clauseActivityDecay", C.c_double), ("quickProf", C.c_int),
                ("initReportCountPerCategory", C.c_int), ("maxPageLockedMemory", C.c_int)]


def available():
    return os.path.exists(REF_SO)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(REF_SO)
        P, I, Q = C.c_void_p, C.c_int, C.c_int64
        sig = {
            "ref_create": (P, [C.POINTER(ref_options)]), "ref_destroy": (None, [P]),
            "ref_gpu_run": (None, [P]), "ref_reduce_db": (None, [P]),
            "ref_set_var_count": (None, [P, I]), "ref_set_cpu_solver_count": (None, [P, I]),
            "ref_add_clause": (Q, [P, I, _IP, I]),
            "ref_add_clauses_bulk": (Q, [P, C.POINTER(Q), _IP, Q]),
            "ref_try_set_solver_values": (I, [P, I, _IP, I]),
            "ref_unset_solver_values": (None, [P, I, _IP, I]),
            "ref_try_send_assignment": (Q, [P, I]),
            "ref_pop_reported_clause": (I, [P, I, C.POINTER(_IP), _IP, C.POINTER(Q)]),
            "ref_get_global_stat": (Q, [P, I]), "ref_get_one_solver_stat": (Q, [P, I, I]),
            "ref_get_last_assig_all_reported": (Q, [P, I]),
            "ref_get_current_assignment": (None, [P, I, C.POINTER(C.c_uint8)]),
            "ref_last_hits": (Q, [P, C.c_void_p, Q]),
        }
        for n, (r, a) in sig.items():
            f = getattr(L, n)
            f.restype, f.argtypes = r, a
        _LIB = L
    return _LIB


def _ints(lits):
    a = np.ascontiguousarray(lits, dtype=np.int32)
    return a, a.ctypes.data_as(_IP), int(a.size)


class RefSharer:
    def __init__(self, blocks=-1, threads=-1, report=-1, min_latency=0, decay=-1.0):
        o = ref_options(blocks, threads, min_latency, 0, decay, 1, report, -1)
        self._L = lib()
        self._h = self._L.ref_create(C.byref(o))

    def close(self):
        if self._h:
            self._L.ref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def gpuRun(self): self._L.ref_gpu_run(self._h)
    def reduceDb(self): self._L.ref_reduce_db(self._h)
    def setVarCount(self, n): self._L.ref_set_var_count(self._h, n)
    def setCpuSolverCount(self, n): self._L.ref_set_cpu_solver_count(self._h, n)

    def addClause(self, s, lits):
        a, p, n = _ints(lits)
        return self._L.ref_add_clause(self._h, s, p, n)

    def addClausesBulk(self, offsets, lits):
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        li = np.ascontiguousarray(lits, dtype=np.int32)
        return self._L.ref_add_clauses_bulk(self._h, off.ctypes.data_as(C.POINTER(C.c_int64)), li.ctypes.data_as(_IP), off.size - 1)

    def trySetSolverValues(self, s, lits):
        a, p, n = _ints(lits)
        return bool(self._L.ref_try_set_solver_values(self._h, s, p, n))

    def unsetSolverValues(self, s, lits):
        a, p, n = _ints(lits)
        self._L.ref_unset_solver_values(self._h, s, p, n)

    def trySendAssignment(self, s): return self._L.ref_try_send_assignment(self._h, s)

    def popReportedClause(self, s):
        lits, count, cid = _IP(), C.c_int(), C.c_int64()
        if not self._L.ref_pop_reported_clause(self._h, s, C.byref(lits), C.byref(count), C.byref(cid)):
            return None
        return [lits[i] for i in range(count.value)], cid.value

    def getGlobalStat(self, st): return self._L.ref_get_global_stat(self._h, int(st))
    def getOneSolverStat(self, s, st): return self._L.ref_get_one_solver_stat(self._h, s, int(st))
    def getLastAssigAllReported(self, s): return self._L.ref_get_last_assig_all_reported(self._h, s)

    def getCurrentAssignment(self, s, nvars):
        buf = np.zeros(nvars, dtype=np.uint8)
        self._L.ref_get_current_assignment(self._h, s, buf.ctypes.data_as(C.POINTER(C.c_uint8)))
        return buf

    def debugLastHits(self):
        n = self._L.ref_last_hits(self._h, None, 0)
        out = np.zeros(n, dtype=HIT_DTYPE)
        if n:
            self._L.ref_last_hits(self._h, out.ctypes.data, n)
        return out
