"""CPU tier: the C-ABI library loads without a GPU and exports every symbol include/*.h declares."""
import os
import re

import gpusharesat_b200.api as api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("gpushare_b200.h",):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(gss_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_loads_and_exports_every_declared_symbol():
    lib = api.load_library()
    decl = declared_symbols()
    assert len(decl) >= 40
    for name in sorted(decl):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    # and the ctypes mirror binds exactly the declared set (nothing undeclared is called)
    assert set(api.SIGNATURES) == decl
    assert lib.gss_version().decode().startswith("gpushare_b200")


def test_options_defaults_match_reference_constructor():
    # GpuClauseSharer.h:49-58
    o = api.gss_options()
    api.load_library().gss_options_default(o)
    assert (o.gpuBlockCountGuideline, o.gpuThreadsPerBlockGuideline, o.minGpuLatencyMicros) == (-1, -1, -1)
    assert o.verbosity == 1 and o.clauseActivityDecay == -1 and o.quickProf == 1
    assert o.initReportCountPerCategory == -1 and o.maxPageLockedMemory == -1


def test_no_cpu_fallback_in_product():
    # the product never references the oracle, and has no host implementation of the check
    src = os.path.join(ROOT, "gpusharesat_b200")
    for dp, _, files in os.walk(src):
        for f in files:
            if f.endswith((".cc", ".cu", ".cuh", ".h", ".py")):
                text = open(os.path.join(dp, f)).read()
                assert "gss_oracle" not in text and "oracle_lib" not in text, f


def test_synthetic_clause_generator_is_thread_count_invariant(monkeypatch):
    """tests/synthlib/gss_synth.h: the clause generator jumps ahead in its one SplitMix64 stream, so
    the database bench.py builds is the same whatever the number of host threads"""
    import numpy as np
    import synth
    sig = synth.sigma(5000, 3)
    out = []
    for threads in ("1", "7"):
        monkeypatch.setenv("GSS_SYNTH_THREADS", threads)
        out.append(synth.clauses(120_000, 5000, 30, sig, 0.9, 5))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    off, lits = out[0]
    assert off[-1] == len(lits) and lits.min() >= 0 and lits.max() < 10000
    assert set(np.diff(off)[:64]) <= {2, 3, 5, 9, 17, 30}  # the Luby-like length mix
