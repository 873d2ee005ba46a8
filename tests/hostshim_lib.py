"""ctypes binding of tests/hostshim/libgss_hostshim.so (host classes of the product, no GPU)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "hostshim", "libgss_hostshim.so")

VARUPDATE = np.dtype([("var", "<i4"), ("def", "<u4"), ("tru", "<u4")])
HIT = np.dtype([("mask", "<u4"), ("solver", "<i4"), ("len", "<i4"), ("idx", "<i4")])


class SolverRunParams(C.Structure):
    _fields_ = [("startVals", C.c_uint32), ("lastMask", C.c_uint32), ("allAggBits", C.c_uint32),
                ("usedAggBits", C.c_uint32), ("updStart", C.c_int32), ("updCount", C.c_int32),
                ("nGroups", C.c_int32), ("pad", C.c_int32),
                ("groupAggBit", C.c_uint32 * 32), ("groupSlotMask", C.c_uint32 * 32)]


def load():
    if not os.path.exists(SO):
        subprocess.check_call(["make", "-C", os.path.dirname(SO)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    L = C.CDLL(SO)
    P, I, Q = C.c_void_p, C.c_int, C.c_int64
    IP = C.POINTER(C.c_int)
    sig = {
        "hs_create": (P, [I, I, C.c_double]), "hs_destroy": (None, [P]),
        "hs_available": (I, [P, I]), "hs_set_var": (None, [P, I, I, I]), "hs_send": (Q, [P, I]),
        "hs_set_host_bumps": (None, [P, I]), "hs_collect": (I, [P, I]), "hs_collect_split": (I, [P]), "hs_collect_take": (I, [P]), "hs_hand_over_views": (C.c_double, [P, C.c_void_p, I, I]), "hs_hand_over_views_shares": (C.c_double, [P, C.c_void_p, I, I, I]), "hs_hand_over_sorted": (C.c_double, [P, C.c_void_p, I]), "hs_get_params": (None, [P, I, C.POINTER(SolverRunParams)]),
        "hs_get_updates": (None, [P, C.c_void_p]), "hs_get_ids": (None, [P, I, C.POINTER(Q), IP]),
        "hs_add_clause": (Q, [P, IP, I]), "hs_drain": (None, [P]), "hs_count": (I, [P, I]),
        "hs_clause_id": (Q, [P, I, I]), "hs_activity": (C.c_float, [P, I, I]), "hs_bump": (None, [P, I, I]),
        "hs_approx_nth_act": (C.c_float, [P, Q]), "hs_reduce_host": (None, [P]),
        "hs_db_clauses": (Q, [P]), "hs_db_length_sum": (Q, [P]), "hs_get_clause": (I, [P, I, I, IP]),
        "hs_clause_was_added": (None, [P, I, Q]), "hs_fill": (None, [P, C.c_void_p, I]),
        "hs_hand_over": (C.c_double, [P, C.c_void_p, I]),
        "hs_add_clauses_bulk": (Q, [P, C.POINTER(Q), IP, Q]),
        "hs_pop": (I, [P, I, IP, IP, C.POINTER(Q)]), "hs_last_all_reported": (Q, [P, I]),
        "hs_solver_stat": (Q, [P, I, I]),
        "hs_shard_directory": (I, [I, I, IP, IP, I, C.POINTER(Q), C.POINTER(Q)]),
    }
    for n, (r, a) in sig.items():
        f = getattr(L, n)
        f.restype, f.argtypes = r, a
    return L


def shard_directory(rank, world, lens, counts):
    """rows (len, count, firstTile, localTiles, ascStart) of a device's directory and the clauses it checks"""
    L = load()
    n = len(lens)
    la, ca = (C.c_int * n)(*lens), (C.c_int * n)(*counts)
    out, tot = (C.c_int64 * (5 * n))(), C.c_int64(0)
    k = L.hs_shard_directory(rank, world, la, ca, n, out, C.byref(tot))
    return [tuple(out[5 * i: 5 * i + 5]) for i in range(k)], tot.value


class Rig:
    def __init__(self, nvars, nsolvers, decay=0.99999):
        self.L = load()
        self.h = self.L.hs_create(nvars, nsolvers, decay)
        self.nsolvers = nsolvers

    def __del__(self):
        try:
            self.L.hs_destroy(self.h)
        except Exception:
            pass

    def set_host_bumps(self, on): self.L.hs_set_host_bumps(self.h, 1 if on else 0)
    def set(self, s, var, val): self.L.hs_set_var(self.h, s, var, val)
    def send(self, s): return self.L.hs_send(self.h, s)
    def available(self, s): return bool(self.L.hs_available(self.h, s))

    def collect(self, full=False):
        n = self.L.hs_collect(self.h, 1 if full else 0)
        upd = np.zeros(n, dtype=VARUPDATE)
        if n:
            self.L.hs_get_updates(self.h, upd.ctypes.data)
        params = []
        for s in range(self.nsolvers):
            p = SolverRunParams()
            self.L.hs_get_params(self.h, s, C.byref(p))
            params.append(p)
        return upd, params

    def collect_split(self):
        """same result as collect(), produced the way the engine collects large batches"""
        n = self.L.hs_collect_split(self.h)
        upd = np.zeros(n, dtype=VARUPDATE)
        if n:
            self.L.hs_get_updates(self.h, upd.ctypes.data)
        params = []
        for s in range(self.nsolvers):
            p = SolverRunParams()
            self.L.hs_get_params(self.h, s, C.byref(p))
            params.append(p)
        return upd, params

    def collect_take(self):
        """same result as collect(), produced the way the direct pipeline collects (buffer swap)"""
        n = self.L.hs_collect_take(self.h)
        upd = np.zeros(n, dtype=VARUPDATE)
        if n:
            self.L.hs_get_updates(self.h, upd.ctypes.data)
        params = []
        for s in range(self.nsolvers):
            p = SolverRunParams()
            self.L.hs_get_params(self.h, s, C.byref(p))
            params.append(p)
        return upd, params

    def hand_over_views(self, hits, parts=1, share_bound=0):
        """share_bound > 0: slices cut like the devices' shares (every length's index range split `parts` ways)"""
        a = np.ascontiguousarray(hits, dtype=HIT)
        return self.L.hs_hand_over_views_shares(self.h, a.ctypes.data, a.size, parts, share_bound)

    def hand_over_sorted(self, hits):
        a = np.ascontiguousarray(hits, dtype=HIT)
        return self.L.hs_hand_over_sorted(self.h, a.ctypes.data, a.size)

    def ids(self, s):
        a, b = C.c_int64(), C.c_int()
        self.L.hs_get_ids(self.h, s, C.byref(a), C.byref(b))
        return a.value, b.value

    def add_clause(self, lits):
        a = np.ascontiguousarray(lits, dtype=np.int32)
        return self.L.hs_add_clause(self.h, a.ctypes.data_as(C.POINTER(C.c_int)), a.size)

    def drain(self): self.L.hs_drain(self.h)
    def count(self, n): return self.L.hs_count(self.h, n)
    def clause_id(self, n, i): return self.L.hs_clause_id(self.h, n, i)
    def activity(self, n, i): return self.L.hs_activity(self.h, n, i)
    def bump(self, n, i): self.L.hs_bump(self.h, n, i)
    def approx_nth_act(self, n): return self.L.hs_approx_nth_act(self.h, n)
    def reduce_host(self): self.L.hs_reduce_host(self.h)
    def db_clauses(self): return self.L.hs_db_clauses(self.h)
    def db_length_sum(self): return self.L.hs_db_length_sum(self.h)

    def get_clause(self, n, i):
        buf = (C.c_int * n)()
        k = self.L.hs_get_clause(self.h, n, i, buf)
        return list(buf[:k])

    def clause_was_added(self, s, cid): self.L.hs_clause_was_added(self.h, s, cid)

    def fill(self, hits):
        a = np.array(hits, dtype=HIT) if len(hits) else np.zeros(0, dtype=HIT)
        self.L.hs_fill(self.h, a.ctypes.data, a.size)

    def hand_over(self, hits):
        a = np.ascontiguousarray(hits, dtype=HIT)
        return self.L.hs_hand_over(self.h, a.ctypes.data, a.size)

    def add_clauses_bulk(self, offsets, lits):
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        li = np.ascontiguousarray(lits, dtype=np.int32)
        return self.L.hs_add_clauses_bulk(self.h, off.ctypes.data_as(C.POINTER(C.c_int64)), li.ctypes.data_as(C.POINTER(C.c_int)), off.size - 1)

    def pop(self, s):
        lits = (C.c_int * 1024)()
        c, i = C.c_int(), C.c_int64()
        if not self.L.hs_pop(self.h, s, lits, C.byref(c), C.byref(i)):
            return None
        return list(lits[:c.value]), i.value

    def pop_all(self, s):
        out = []
        while True:
            r = self.pop(s)
            if r is None:
                return out
            out.append(r)

    def last_all_reported(self, s): return self.L.hs_last_all_reported(self.h, s)
    def stat(self, s, k): return self.L.hs_solver_stat(self.h, s, k)


class DeviceModel:
    """numpy statement of what k_apply_updates / k_collapse do to the tables (used to check the
    host-produced run parameters against the reference's golden words)"""

    def __init__(self, nvars, nsolvers):
        self.def_ = np.zeros((nsolvers, nvars), dtype=np.uint32)
        self.tru = np.zeros((nsolvers, nvars), dtype=np.uint32)
        self.can_true = np.zeros(nvars, dtype=np.uint32)
        self.can_false = np.zeros(nvars, dtype=np.uint32)
        self.can_undef = np.full(nvars, 0xFFFFFFFF, dtype=np.uint32)

    @staticmethod
    def _merge(arr, v, mask, bits):
        arr[v] = np.uint32((int(arr[v]) & ~mask) | bits)

    def apply(self, upd, params):
        for s, p in enumerate(params):
            for u in upd[p.updStart:p.updStart + p.updCount]:
                v, d, t = int(u["var"]), int(u["def"]), int(u["tru"])
                self.def_[s, v], self.tru[s, v] = d, t
                T = F = U = 0
                for g in range(p.nGroups):
                    m, bit = p.groupSlotMask[g], p.groupAggBit[g]
                    if t & d & m: T |= bit
                    if ~t & d & m: F |= bit
                    if ~d & m & 0xFFFFFFFF: U |= bit
                self._merge(self.can_true, v, p.usedAggBits, T)
                self._merge(self.can_false, v, p.usedAggBits, F)
                self._merge(self.can_undef, v, p.usedAggBits, U)

    def collapse(self, upd, params):
        for s, p in enumerate(params):
            for u in upd[p.updStart:p.updStart + p.updCount]:
                v = int(u["var"])
                d = 0xFFFFFFFF if int(u["def"]) & p.lastMask else 0
                t = 0xFFFFFFFF if int(u["tru"]) & p.lastMask else 0
                self.def_[s, v], self.tru[s, v] = d, t
                a = p.allAggBits
                self._merge(self.can_true, v, a, a if (d & t) else 0)
                self._merge(self.can_false, v, a, a if (d & ~t) else 0)
                self._merge(self.can_undef, v, a, 0 if d else a)
