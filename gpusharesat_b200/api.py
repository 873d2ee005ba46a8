"""ctypes mirror of ``GpuShare::GpuClauseSharer`` (reference gpuShareLib/GpuClauseSharer.h:73-161)
on top of the C ABI in ``include/gpushare_b200.h``.  Same method names, argument meaning and
return conventions as the reference class, so tests read like the reference's own.

No CPU fallback: :func:`load_library` raises if ``libgpushare_b200.so`` has not been built
(``python -c "import __graft_entry__ as g; g.build()"``), and ``GpuClauseSharer()`` exits the
process (like the reference) if no CUDA device is usable.
"""
import ctypes as C
import enum
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def library_path():
    return os.path.join(_HERE, "libgpushare_b200.so")


class gss_options(C.Structure):
    # include/gpushare_b200.h: gss_options  (GpuClauseSharer.h:25-59)
    _fields_ = [("gpuBlockCountGuideline", C.c_int), ("gpuThreadsPerBlockGuideline", C.c_int),
                ("minGpuLatencyMicros", C.c_int), ("verbosity", C.c_int),
                ("clauseActivityDecay", C.c_double), ("quickProf", C.c_int),
                ("initReportCountPerCategory", C.c_int), ("maxPageLockedMemory", C.c_int)]


class gss_hit(C.Structure):
    _fields_ = [("clause_id", C.c_int64), ("solver_id", C.c_int32), ("mask", C.c_uint32)]


HIT_DTYPE = np.dtype([("clause_id", "<i8"), ("solver_id", "<i4"), ("mask", "<u4")])

RAW_HIT_DTYPE = np.dtype([("mask", "<u4"), ("solver", "<i4"), ("len", "<i4"), ("idx", "<i4")])

_LOGFN = C.CFUNCTYPE(None, C.c_char_p, C.c_void_p)

_P = C.c_void_p
_I = C.c_int
_L = C.c_int64
_IP = C.POINTER(C.c_int)

# name -> (restype, argtypes): every symbol include/gpushare_b200.h declares
SIGNATURES = {
    "gss_options_default": (None, [C.POINTER(gss_options)]),
    "gss_create": (_P, [C.POINTER(gss_options), _LOGFN, C.c_void_p]),
    "gss_destroy": (None, [_P]),
    "gss_gpu_run": (None, [_P]),
    "gss_reduce_db": (None, [_P]),
    "gss_get_added_clause_count": (_L, [_P]),
    "gss_get_added_clause_count_at_last_reduce_db": (_L, [_P]),
    "gss_has_run_out_of_gpu_memory_once": (_I, [_P]),
    "gss_get_gpu_mem_info": (None, [_P, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "gss_get_global_stat_count": (_I, [_P]),
    "gss_get_global_stat": (_L, [_P, _I]),
    "gss_get_global_stat_name": (C.c_char_p, [_P, _I]),
    "gss_write_clauses_in_cnf": (None, [_P, C.c_void_p]),
    "gss_set_var_count": (None, [_P, _I]),
    "gss_set_cpu_solver_count": (None, [_P, _I]),
    "gss_add_clause": (_L, [_P, _I, _IP, _I]),
    "gss_try_set_solver_values": (_I, [_P, _I, _IP, _I]),
    "gss_unset_solver_values": (None, [_P, _I, _IP, _I]),
    "gss_try_send_assignment": (_L, [_P, _I]),
    "gss_pop_reported_clause": (_I, [_P, _I, C.POINTER(_IP), _IP, C.POINTER(_L)]),
    "gss_get_last_assig_all_reported": (_L, [_P, _I]),
    "gss_get_current_assignment": (None, [_P, _I, C.POINTER(C.c_uint8)]),
    "gss_get_one_solver_stat_count": (_I, [_P]),
    "gss_get_one_solver_stat": (_L, [_P, _I, _I]),
    "gss_get_one_solver_stat_name": (C.c_char_p, [_P, _I]),
    "gss_debug_last_hits": (_L, [_P, C.c_void_p, _L]),
    "gss_add_clauses_bulk": (_L, [_P, C.POINTER(_L), _IP, _L]),
    "gss_set_max_clause_len": (None, [_P, _I]),
    "gss_debug_set_dense": (None, [_P, _I]),
    "gss_debug_time_check": (C.c_double, [_P, _I, _I]),
    "gss_debug_last_run_times": (_I, [_P, C.POINTER(C.c_double)]),
    "gss_debug_host_phases": (None, [_P, C.POINTER(C.c_double)]),
    "gss_debug_filter_variants": (_I, []),
    "gss_debug_filter_variant_name": (C.c_char_p, [_I]),
    "gss_debug_set_filter_variant": (None, [_I]),
    "gss_debug_exact_variants": (_I, []),
    "gss_debug_exact_variant_name": (C.c_char_p, [_I]),
    "gss_debug_set_exact_variant": (None, [_I]),
    "gss_debug_lop3_peak": (C.c_double, [_P]),
    "gss_debug_last_run_bytes": (None, [_P, C.POINTER(_L), C.POINTER(_L)]),
    "gss_debug_kernel_launches": (_L, [_P]),
    "gss_debug_db_size": (None, [_P, C.POINTER(_L), C.POINTER(_L)]),
    "gss_debug_db_order": (None, [_P, C.POINTER(_L), C.POINTER(_L)]),
    "gss_set_shard": (None, [_P, _I, _I]),
    "gss_mgpu_collect": (_I, [_P, C.POINTER(C.c_void_p), C.POINTER(_L), C.POINTER(C.c_void_p), C.POINTER(_L)]),
    "gss_mgpu_run": (None, [_P, C.c_void_p, _L, C.c_void_p, _L, _I]),
    "gss_mgpu_wait": (_L, [_P, C.POINTER(C.c_void_p)]),
    "gss_mgpu_collect_to": (_L, [_P, C.c_void_p, _L]),
    "gss_mgpu_run_payload": (_I, [_P, C.c_void_p, _L]),
    "gss_mgpu_hits_to_device": (_L, [_P, C.c_void_p, _L]),
    "gss_mgpu_enqueue_payload": (_I, [_P, C.c_void_p, _L]),
    "gss_mgpu_redo_payload": (None, [_P, C.c_void_p, _L]),
    "gss_mgpu_enqueue_result": (_L, [_P, C.c_void_p, _L]),
    "gss_mgpu_finish": (_I, [_P]),
    "gss_mgpu_import_gathered": (None, [_P, C.c_void_p, _I, _L, C.POINTER(_L)]),
    "gss_peer_init": (_L, [_P, _I, _I, _L, _L, C.c_void_p, _L]),
    "gss_peer_connect": (None, [_P, C.c_void_p, _L]),
    "gss_peer_enqueue": (_I, [_P]),
    "gss_peer_finish": (_L, [_P]),
    "gss_debug_set_peer_records": (None, [_P, _I]),
    "gss_set_stream": (None, [_P, C.c_void_p]),
    "gss_mgpu_import": (None, [_P, C.c_void_p, _L]),
    "gss_version": (C.c_char_p, []),
}


def load_library(path=None):
    """Load libgpushare_b200.so and bind every declared symbol.  Raises if it is missing."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or library_path()
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: the CUDA extension has not been built (run __graft_entry__.build()). "
            "gpusharesat_b200 has no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _LIB = lib
    return lib


def mkLit(var, sign=False):
    """reference mkLit (gpuShareLib/SolverTypes.h:56): 2*var + sign, sign True = negated"""
    return 2 * int(var) + (1 if sign else 0)


class GlobalStats(enum.IntEnum):  # gpuShareLib/GlobalStats.h:23-38
    gpuClauses = 0
    gpuClauseLengthSum = 1
    gpuClausesAdded = 2
    gpuRuns = 3
    clauseTestsOnGroups = 4
    clauseTestsOnAssigs = 5
    totalAssigClauseTested = 6
    gpuReduceDbs = 7
    gpuReports = 8
    timeSpentTestingClauses = 9
    timeSpentFillingAssigs = 10
    timeSpentFillingReported = 11
    timeSpentReduceGpuDb = 12


class OneSolverStats(enum.IntEnum):  # gpuShareLib/OneSolverStats.h:24-29
    varUpdatesSentToGpu = 0
    assigsSentToGpu = 1
    failuresToFindAssig = 2
    reportedClauses = 3
    reportedClausesUnit = 4
    reportedClausesBinary = 5


class GpuClauseSharerOptions:
    """reference GpuClauseSharerOptions (GpuClauseSharer.h:25-59); -1 = default"""

    def __init__(self, **kw):
        self.gpuBlockCountGuideline = -1
        self.gpuThreadsPerBlockGuideline = -1
        self.minGpuLatencyMicros = -1
        self.verbosity = 0
        self.clauseActivityDecay = -1.0
        self.quickProf = True
        self.initReportCountPerCategory = -1
        self.maxPageLockedMemory = -1
        for k, v in kw.items():
            if not hasattr(self, k):
                raise TypeError(f"unknown option {k}")
            setattr(self, k, v)

    def _c(self):
        o = gss_options()
        o.gpuBlockCountGuideline = self.gpuBlockCountGuideline
        o.gpuThreadsPerBlockGuideline = self.gpuThreadsPerBlockGuideline
        o.minGpuLatencyMicros = self.minGpuLatencyMicros
        o.verbosity = self.verbosity
        o.clauseActivityDecay = self.clauseActivityDecay
        o.quickProf = 1 if self.quickProf else 0
        o.initReportCountPerCategory = self.initReportCountPerCategory
        o.maxPageLockedMemory = self.maxPageLockedMemory
        return o


def _ints(lits):
    a = np.ascontiguousarray(lits, dtype=np.int32)
    return a, a.ctypes.data_as(_IP), int(a.size)


class GpuClauseSharer:
    """Mirror of the reference's abstract class; see include/gpushare_b200.h for the line map."""

    def __init__(self, opts=None, logFunc=None, lib=None):
        self._lib = lib or load_library()
        self._opts = opts or GpuClauseSharerOptions()
        self._logfn = _LOGFN(lambda msg, ctx: logFunc(msg.decode())) if logFunc else C.cast(None, _LOGFN)
        o = self._opts._c()
        self._h = self._lib.gss_create(C.byref(o), self._logfn, None)
        if not self._h:
            raise RuntimeError("gss_create failed")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gss_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- GPU thread ----
    def gpuRun(self):
        self._lib.gss_gpu_run(self._h)

    def reduceDb(self):
        self._lib.gss_reduce_db(self._h)

    def getAddedClauseCount(self):
        return self._lib.gss_get_added_clause_count(self._h)

    def getAddedClauseCountAtLastReduceDb(self):
        return self._lib.gss_get_added_clause_count_at_last_reduce_db(self._h)

    def hasRunOutOfGpuMemoryOnce(self):
        return bool(self._lib.gss_has_run_out_of_gpu_memory_once(self._h))

    def getGpuMemInfo(self):
        f, t = C.c_size_t(), C.c_size_t()
        self._lib.gss_get_gpu_mem_info(self._h, C.byref(f), C.byref(t))
        return f.value, t.value

    def getGlobalStatCount(self):
        return self._lib.gss_get_global_stat_count(self._h)

    def getGlobalStat(self, stat):
        return self._lib.gss_get_global_stat(self._h, int(stat))

    def getGlobalStatName(self, stat):
        return self._lib.gss_get_global_stat_name(self._h, int(stat)).decode()

    def writeClausesInCnf(self, path):
        libc = C.CDLL(None)
        libc.fopen.restype = C.c_void_p
        libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
        libc.fclose.argtypes = [C.c_void_p]
        f = libc.fopen(path.encode(), b"w")
        if not f:
            raise OSError(f"cannot open {path}")
        self._lib.gss_write_clauses_in_cnf(self._h, f)
        libc.fclose(f)

    def setVarCount(self, n):
        self._lib.gss_set_var_count(self._h, int(n))

    def setCpuSolverCount(self, n):
        self._lib.gss_set_cpu_solver_count(self._h, int(n))

    # ---- any thread ----
    def addClause(self, solverId, lits):
        a, p, n = _ints(lits)
        return self._lib.gss_add_clause(self._h, int(solverId), p, n)

    # ---- solver threads ----
    def trySetSolverValues(self, solverId, lits):
        a, p, n = _ints(lits)
        return bool(self._lib.gss_try_set_solver_values(self._h, int(solverId), p, n))

    def unsetSolverValues(self, solverId, lits):
        a, p, n = _ints(lits)
        self._lib.gss_unset_solver_values(self._h, int(solverId), p, n)

    def trySendAssignment(self, solverId):
        return self._lib.gss_try_send_assignment(self._h, int(solverId))

    def popReportedClause(self, solverId):
        """returns (lits list, gpuClauseId) or None"""
        lits, count, cid = _IP(), C.c_int(), C.c_int64()
        if not self._lib.gss_pop_reported_clause(self._h, int(solverId), C.byref(lits), C.byref(count), C.byref(cid)):
            return None
        return [lits[i] for i in range(count.value)], cid.value

    def getLastAssigAllReported(self, solverId):
        return self._lib.gss_get_last_assig_all_reported(self._h, int(solverId))

    def getCurrentAssignment(self, solverId, varCount):
        buf = np.zeros(varCount, dtype=np.uint8)
        self._lib.gss_get_current_assignment(self._h, int(solverId), buf.ctypes.data_as(C.POINTER(C.c_uint8)))
        return buf

    def getOneSolverStatCount(self):
        return self._lib.gss_get_one_solver_stat_count(self._h)

    def getOneSolverStat(self, solverId, stat):
        return self._lib.gss_get_one_solver_stat(self._h, int(solverId), int(stat))

    def getOneSolverStatName(self, stat):
        return self._lib.gss_get_one_solver_stat_name(self._h, int(stat)).decode()

    # ---- parity / bench hooks (not in GpuClauseSharer.h) ----
    def debugLastHits(self):
        n = self._lib.gss_debug_last_hits(self._h, None, 0)
        out = np.zeros(n, dtype=HIT_DTYPE)
        if n:
            self._lib.gss_debug_last_hits(self._h, out.ctypes.data, n)
        return out

    def addClausesBulk(self, offsets, lits):
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        li = np.ascontiguousarray(lits, dtype=np.int32)
        return self._lib.gss_add_clauses_bulk(self._h, off.ctypes.data_as(C.POINTER(_L)), li.ctypes.data_as(_IP), int(off.size - 1))

    def setMaxClauseLen(self, n):
        self._lib.gss_set_max_clause_len(self._h, int(n))

    def debugSetDense(self, dense):
        self._lib.gss_debug_set_dense(self._h, 1 if dense else 0)

    def debugTimeCheck(self, iters=10, dense=False, filter_only=False, mode=None):
        """mode: 0 production check, 1 dense, 2 k_filter, 3 k_exact, 4 k_apply_updates, 5 k_collapse, 6 k_emit"""
        if mode is None:
            mode = 2 if filter_only else (1 if dense else 0)
        return self._lib.gss_debug_time_check(self._h, int(iters), int(mode))

    def debugHostPhases(self):
        t = (C.c_double * 6)()
        self._lib.gss_debug_host_phases(self._h, t)
        return list(t)

    def debugFilterVariants(self):
        return [self._lib.gss_debug_filter_variant_name(v).decode() for v in range(self._lib.gss_debug_filter_variants())]

    def debugSetFilterVariant(self, v):
        self._lib.gss_debug_set_filter_variant(int(v))

    def debugExactVariants(self):
        return [self._lib.gss_debug_exact_variant_name(v).decode() for v in range(self._lib.gss_debug_exact_variants())]

    def debugSetExactVariant(self, v):
        self._lib.gss_debug_set_exact_variant(int(v))

    def debugLastRunTimes(self):
        t = (C.c_double * 4)()
        if not self._lib.gss_debug_last_run_times(self._h, t):
            return None
        return list(t)

    def debugLop3Peak(self):
        return self._lib.gss_debug_lop3_peak(self._h)

    def debugLastRunBytes(self):
        a, b = C.c_int64(), C.c_int64()
        self._lib.gss_debug_last_run_bytes(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def debugKernelLaunches(self):
        return self._lib.gss_debug_kernel_launches(self._h)

    # ---- multi-GPU (include/gpushare_b200.h) ----
    def setShard(self, rank, world):
        self._lib.gss_set_shard(self._h, int(rank), int(world))

    def mgpuCollect(self):
        """rank 0: returns (rebuild, params_ptr, params_bytes, updates_ptr, n_updates) or None"""
        pp, pb, up, un = C.c_void_p(), C.c_int64(), C.c_void_p(), C.c_int64()
        r = self._lib.gss_mgpu_collect(self._h, C.byref(pp), C.byref(pb), C.byref(up), C.byref(un))
        if r < 0:
            return None
        return r, pp.value, pb.value, up.value or 0, un.value

    def mgpuRun(self, params_ptr, params_bytes, updates_ptr, n_updates, rebuild):
        self._lib.gss_mgpu_run(self._h, params_ptr, params_bytes, updates_ptr, n_updates, int(rebuild))

    def mgpuCollectTo(self, dev_ptr, cap_bytes):
        return self._lib.gss_mgpu_collect_to(self._h, dev_ptr, int(cap_bytes))

    def mgpuRunPayload(self, dev_ptr, nbytes):
        return self._lib.gss_mgpu_run_payload(self._h, dev_ptr, int(nbytes))

    def mgpuHitsToDevice(self, dev_ptr, cap_records):
        return self._lib.gss_mgpu_hits_to_device(self._h, dev_ptr, int(cap_records))

    def mgpuEnqueuePayload(self, dev_ptr, valid_bytes):
        return self._lib.gss_mgpu_enqueue_payload(self._h, dev_ptr, int(valid_bytes))

    def mgpuRedoPayload(self, dev_ptr, total_bytes):
        self._lib.gss_mgpu_redo_payload(self._h, dev_ptr, int(total_bytes))

    def mgpuEnqueueResult(self, dev_ptr, cap_records):
        return self._lib.gss_mgpu_enqueue_result(self._h, dev_ptr, int(cap_records))

    def mgpuImportGathered(self, dev_ptr, world, slot_bytes, counts):
        c = np.ascontiguousarray(counts, dtype=np.int64)
        self._lib.gss_mgpu_import_gathered(self._h, dev_ptr, int(world), int(slot_bytes), c.ctypes.data_as(C.POINTER(_L)))

    def peerInit(self, rank, world, payload_cap, slot_hits):
        buf = C.create_string_buffer(128)
        n = self._lib.gss_peer_init(self._h, rank, world, int(payload_cap), int(slot_hits), buf, 128)
        return buf.raw[:n]

    def peerConnect(self, blobs):
        """blobs: list of the ranks' peerInit() results, in rank order"""
        stride = len(blobs[0])
        raw = b"".join(blobs)
        self._lib.gss_peer_connect(self._h, C.c_char_p(raw), stride)

    def peerEnqueue(self):
        return self._lib.gss_peer_enqueue(self._h)

    def peerFinish(self):
        return self._lib.gss_peer_finish(self._h)

    def debugSetPeerRecords(self, on):
        self._lib.gss_debug_set_peer_records(self._h, 1 if on else 0)

    def mgpuFinish(self):
        return self._lib.gss_mgpu_finish(self._h)

    def mgpuWaitCount(self):
        """wait for the run; returns this rank's hit count (hits stay on the device / in the library)"""
        return self._lib.gss_mgpu_wait(self._h, None)

    def setStream(self, cuda_stream):
        self._lib.gss_set_stream(self._h, C.c_void_p(cuda_stream))

    def mgpuWait(self):
        """returns this rank's hits as a numpy array (mask, solver, len, idx) -- a copy"""
        p = C.c_void_p()
        n = self._lib.gss_mgpu_wait(self._h, C.byref(p))
        out = np.zeros(n, dtype=RAW_HIT_DTYPE)
        if n:
            C.memmove(out.ctypes.data, p.value, n * RAW_HIT_DTYPE.itemsize)
        return out

    def mgpuImport(self, hits):
        a = np.ascontiguousarray(hits, dtype=RAW_HIT_DTYPE)
        self._lib.gss_mgpu_import(self._h, a.ctypes.data, int(a.size))

    def debugDbOrder(self):
        """(clauses behind the sorted part of their arena, device-side re-sorts so far)"""
        a, b = C.c_int64(), C.c_int64()
        self._lib.gss_debug_db_order(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def debugDbSize(self):
        a, b = C.c_int64(), C.c_int64()
        self._lib.gss_debug_db_size(self._h, C.byref(a), C.byref(b))
        return a.value, b.value
