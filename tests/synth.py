"""Synthetic inputs of the BASELINE.json shapes via tests/synthlib/libgss_synth.so (gss_synth.h): a
host-only library of its own, so generating inputs (e.g. for bench.py --impl reference) never loads
the product library.  Input generation only."""
import ctypes as C
import os
import subprocess

import numpy as np

_IP = C.POINTER(C.c_int)
_U8 = C.POINTER(C.c_uint8)
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "synthlib")
_LIB = None


def load_library():
    global _LIB
    if _LIB is None:
        so = os.path.join(_DIR, "libgss_synth.so")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(_DIR, "synth.cc")):
            subprocess.check_call(["make", "-C", _DIR], stdout=subprocess.DEVNULL)
        L = C.CDLL(so)
        _L, _I, _P = C.c_int64, C.c_int, C.c_void_p
        sig = {
            "gss_synth_total_lits": (_L, [_L, _I]),
            "gss_synth_sigma": (None, [_I, C.c_uint64, _U8]),
            "gss_synth_clauses": (None, [_L, _I, _I, _U8, C.c_double, C.c_uint64, C.POINTER(_L), _IP]),
            "gss_synth_stream_create": (_P, [_I, _U8, C.c_double, C.c_double, C.c_uint64]),
            "gss_synth_stream_destroy": (None, [_P]),
            "gss_synth_stream_values": (_U8, [_P]),
            "gss_synth_stream_next": (None, [_P, _IP, _IP, _IP, _IP]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _LIB = L
    return _LIB


def sigma(nvars, seed):
    s = np.zeros(nvars, dtype=np.uint8)
    load_library().gss_synth_sigma(nvars, seed, s.ctypes.data_as(_U8))
    return s


def clauses(nclauses, nvars, max_len=30, sig=None, p_agree=0.98, seed=1):
    L = load_library()
    total = L.gss_synth_total_lits(nclauses, max_len)
    offsets = np.zeros(nclauses + 1, dtype=np.int64)
    lits = np.zeros(total, dtype=np.int32)
    sp = sig.ctypes.data_as(_U8) if sig is not None else C.cast(None, _U8)
    L.gss_synth_clauses(nclauses, nvars, max_len, sp, p_agree, seed, offsets.ctypes.data_as(C.POINTER(C.c_int64)),
                        lits.ctypes.data_as(_IP))
    return offsets, lits


class Stream:
    """one solver's assignment stream (delta per step)"""

    def __init__(self, nvars, sig, p_undef=0.01, churn=0.01, seed=1):
        self.L = load_library()
        self.nvars = nvars
        self.h = self.L.gss_synth_stream_create(nvars, sig.ctypes.data_as(_U8), p_undef, churn, seed)
        self._set = np.zeros(nvars, dtype=np.int32)
        self._unset = np.zeros(nvars, dtype=np.int32)

    def next(self):
        ns, nu = C.c_int(), C.c_int()
        self.L.gss_synth_stream_next(self.h, self._set.ctypes.data_as(_IP), C.byref(ns),
                                     self._unset.ctypes.data_as(_IP), C.byref(nu))
        return self._set[:ns.value], self._unset[:nu.value]

    def values(self):
        p = self.L.gss_synth_stream_values(self.h)
        return np.ctypeslib.as_array(p, shape=(self.nvars,)).copy()

    def __del__(self):
        try:
            self.L.gss_synth_stream_destroy(self.h)
        except Exception:
            pass


def pack_slots(snapshots):
    """list of <=32 value arrays (0 true / 1 false / 2 undef) -> (def, tru, start) words"""
    nvars = snapshots[0].size
    d = np.zeros(nvars, dtype=np.uint32)
    t = np.zeros(nvars, dtype=np.uint32)
    for p, v in enumerate(snapshots):
        bit = np.uint32(1 << p)
        d[v != 2] |= bit
        t[v == 0] |= bit
    start = np.uint32((1 << len(snapshots)) - 1) if len(snapshots) < 32 else np.uint32(0xFFFFFFFF)
    return d, t, start
