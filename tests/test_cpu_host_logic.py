"""CPU tier: the product's host logic (slot state machine, run parameters, hand-over rules,
reduceDb thresholds) through tests/hostshim, against the reference's golden values."""
import numpy as np
import pytest

from hostshim_lib import HIT, DeviceModel, Rig

TRUE, FALSE, UNDEF = 0, 1, 2
ALL = 0xFFFFFFFF


def test_two_solvers_one_assignment_each():
    # GpuSolverTest.cu:78-121 testAssigsTwoSolvers
    r = Rig(2, 2)
    r.set(0, 0, FALSE); r.set(0, 1, UNDEF); assert r.send(0) == 0
    r.set(1, 0, UNDEF); r.set(1, 1, TRUE); assert r.send(1) == 0
    upd, params = r.collect()
    dev = DeviceModel(2, 2)
    dev.apply(upd, params)
    assert dev.def_[0, 0] & 1 == 1 and dev.tru[0, 0] & 1 == 0 and dev.def_[0, 1] & 1 == 0
    assert dev.def_[1, 0] & 1 == 0 and dev.def_[1, 1] & 1 == 1 and dev.tru[1, 1] & 1 == 1
    assert params[0].startVals == 1 and params[1].startVals == 1


def test_two_assignments_value_inheritance():
    # GpuSolverTest.cu:123-161 testAssigsTwoAssignments
    r = Rig(2, 1)
    r.set(0, 0, TRUE); r.set(0, 1, FALSE); r.send(0)
    r.set(0, 0, UNDEF); r.send(0)
    upd, params = r.collect()
    dev = DeviceModel(2, 1)
    dev.apply(upd, params)
    assert dev.def_[0, 0] & 1 and dev.tru[0, 0] & 1 and dev.def_[0, 1] & 1 and not dev.tru[0, 1] & 1
    assert not dev.def_[0, 0] & 2
    assert dev.def_[0, 1] & 2 and not dev.tru[0, 1] & 2   # still false in the second slot
    assert params[0].startVals == 3 and params[0].lastMask == 4  # in-progress slot 2 mirrors lastVarVal


def test_32_slots_then_reuse_after_collapse():
    # GpuSolverTest.cu:163-211 testManyAssignments
    r = Rig(32, 1)
    for i in range(32):
        assert r.available(0)
        r.set(0, i, TRUE if i % 2 == 0 else FALSE)
        assert r.send(0) == i
    assert not r.available(0)
    upd, params = r.collect()
    assert params[0].startVals == ALL and params[0].lastMask == 1 << 31
    assert r.ids(0) == (0, 32)
    dev = DeviceModel(32, 1)
    dev.apply(upd, params)
    dev.collapse(upd, params)
    assert r.available(0)
    r.set(0, 7, UNDEF); assert r.send(0) == 32
    upd, params = r.collect()
    dev.apply(upd, params)
    assert params[0].startVals == 1 and r.ids(0) == (32, 1)
    for i in range(32):
        if i == 7:
            assert dev.def_[0, i] & 1 == 0
        else:
            assert dev.def_[0, i] & 1 == 1 and bool(dev.tru[0, i] & 1) == (i % 2 == 0)


def test_aggregate_words_golden():
    # GpuSolverTest.cu:213-243 testAssigAggregates: 2 solvers x 32 slots on one variable
    r = Rig(1, 2)
    for s in range(2):
        for i in range(32):
            r.set(s, 0, (TRUE if s == 0 else FALSE) if i % 2 == 0 else UNDEF)
            r.send(s)
    upd, params = r.collect()
    assert params[0].usedAggBits | params[1].usedAggBits == ALL
    assert params[0].allAggBits == 0x0000FFFF and params[1].allAggBits == 0xFFFF0000
    dev = DeviceModel(1, 2)
    dev.apply(upd, params)
    assert dev.can_true[0] == 0x0000FFFF
    assert dev.can_false[0] == 0xFFFF0000
    assert dev.can_undef[0] == ALL


def test_aggregate_bit_partition_like_reference():
    # Assigs.cu:402-426: 32 bits over n solvers, the first 32 % n get one more
    for n in (1, 2, 3, 5, 7, 32):
        r = Rig(1, n)
        for s in range(n):
            r.set(s, 0, TRUE); r.send(s)
        _, params = r.collect()
        sizes = [bin(p.allAggBits).count("1") for p in params]
        assert sum(sizes) == 32 and sizes == sorted(sizes, reverse=True)
        assert max(sizes) - min(sizes) <= 1
        union = 0
        for p in params:
            assert union & p.allAggBits == 0
            union |= p.allAggBits
            assert p.nGroups == 1  # one frozen slot -> one aggregate bit in use
    # more than 32 solvers: every group of 32 has its own aggregate word (the reference gives
    # solvers 32.. no bit at all, Assigs.cu:409-425)
    r = Rig(1, 40)
    for s in range(40):
        r.send(s)
    _, params = r.collect()
    assert all(p.allAggBits != 0 for p in params)
    assert params[32].allAggBits & 1 and sum(bin(p.allAggBits).count("1") for p in params[32:]) == 32


def test_slot_groups_cover_frozen_slots():
    # Assigs.cu:263-284: k = min(bits, frozen) groups, sizes differ by at most one, ids in order
    r = Rig(1, 3)  # solver 0 owns 11 aggregate bits
    for i in range(25):
        r.send(0)
    _, params = r.collect()
    p = params[0]
    assert p.nGroups == 11 and p.startVals == (1 << 25) - 1
    sizes = [bin(p.groupSlotMask[g]).count("1") for g in range(11)]
    assert sizes == [3, 3, 3] + [2] * 8
    u = 0
    for g in range(11):
        assert u & p.groupSlotMask[g] == 0
        u |= p.groupSlotMask[g]
    assert u == p.startVals


def test_full_rebuild_lists_every_variable():
    r = Rig(5, 1)
    r.set(0, 1, TRUE); r.set(0, 3, FALSE); r.send(0)
    r.collect()
    r.set(0, 3, UNDEF); r.send(0)
    upd, params = r.collect(full=True)
    assert upd["var"].tolist() == [0, 1, 2, 3, 4]
    # untouched variables carry their current value in every slot; variable 3 was unset after the
    # previous batch had shipped, so the change covers all 32 slots
    assert upd["def"].tolist() == [0, ALL, 0, 0, 0] and int(upd["tru"][1]) == ALL
    assert params[0].startVals == 2 and params[0].updCount == 5


def _run(r, hits):
    r.collect()
    r.fill(hits)


def test_reported_counts_and_progress_marker():
    r = Rig(3, 2)
    ids = [r.add_clause([2 * v]) for v in range(3)]
    assert ids == [0, 1, 2]
    r.drain()
    assert r.count(1) == 3 and r.db_clauses() == 3 and r.db_length_sum() == 3
    r.send(0); r.send(0); r.send(1)
    _run(r, [(3, 0, 1, 0), (1, 0, 1, 1), (2, 1, 1, 2)])
    assert r.last_all_reported(0) == 0
    assert [x[1] for x in r.pop_all(0)] == [0, 1]
    assert r.pop_all(1) == [([4], 2)]
    assert r.last_all_reported(0) == 2 and r.last_all_reported(1) == 1
    assert r.stat(0, 3) == 2 and r.stat(0, 4) == 2  # reportedClauses, reportedClausesUnit


def test_no_double_import_across_runs_and_reimport_later():
    # GpuSolverTest.cu:540-563 and :566-603 / Reported.cu:105-158
    r = Rig(2, 1)
    r.add_clause([1, 2]); r.drain()
    r.send(0)
    _run(r, [(1, 0, 2, 0)])          # run A: assignment 0
    r.send(0)                        # assignment 1 sent before the solver saw the report
    assert len(r.pop_all(0)) == 1
    _run(r, [(2, 0, 2, 0)])          # run B: assignment 1 did not know the clause -> suppressed
    assert r.pop_all(0) == []
    r.send(0)                        # assignment 2 knows it; if it fires again the solver deleted it
    _run(r, [(4, 0, 2, 0)])
    assert len(r.pop_all(0)) == 1
    assert r.stat(0, 3) == 2 and r.stat(0, 5) == 2  # reportedClauses, reportedClausesBinary


def test_exporter_echo_suppressed_until_known():
    # Reported.cu:97-103 clauseWasAdded
    r = Rig(2, 2)
    r.send(0); r.send(1)
    cid = r.add_clause([1, 3]); r.clause_was_added(0, cid); r.drain()
    _run(r, [(1, 0, 2, 0), (1, 1, 2, 0)])
    assert r.pop_all(0) == [] and len(r.pop_all(1)) == 1


def _dup_scenario():
    r = Rig(4, 1)
    for v in range(3):
        r.add_clause([2 * v])
    r.drain()
    r.send(0)
    _run(r, [(1, 0, 1, 1)])              # clause 1 reported for assignment 0
    r.send(0)
    assert len(r.pop_all(0)) == 1
    _run(r, [(2, 0, 1, 0), (2, 0, 1, 1), (2, 0, 1, 2)])   # batch: clause 0, then duplicate 1, then 2
    return [g[1] for g in r.pop_all(0)]


def test_duplicate_is_skipped_not_the_rest_of_the_batch(monkeypatch):
    # Reported.cu:113-129: in the reference a re-reported clause falls through to the end-of-batch
    # block and the rest of the batch (clause 2 here) is silently lost.  Deliberate difference:
    # only the duplicate is skipped; GPUSHARE_REFERENCE_DUP_QUIRK=1 restores the reference behaviour.
    monkeypatch.delenv("GPUSHARE_REFERENCE_DUP_QUIRK", raising=False)
    assert _dup_scenario() == [0, 2]
    monkeypatch.setenv("GPUSHARE_REFERENCE_DUP_QUIRK", "1")
    assert _dup_scenario() == [0]


def test_activity_decay_bump_and_reduce():
    # ClauseActivityLbdTest.cu:59-275 / Clauses.cu:200-237,426-465 (activity-only policy)
    r = Rig(10, 1, decay=0.5)
    for i in range(8):
        r.add_clause([2 * i, 2 * i + 2, 2 * ((i + 2) % 10)])   # length 3
    r.add_clause([0, 2]); r.add_clause([4])
    r.drain()
    acts = [r.activity(3, i) for i in range(8)]
    assert acts == [2.0 ** (i + 1) for i in range(8)]          # increment doubles per added clause
    r.bump(3, 0)
    assert r.activity(3, 0) == 2.0 + 2.0 ** 10
    thr = r.approx_nth_act(5)
    # activities: 4 8 16 32 [64] 128 256 512(len 2) 1024(len 1) 1026(bumped); the 5th smallest is 64 and
    # the log-scale bucket is rounded UP (Clauses.cu:521), so the threshold lies just above it
    assert 64.0 < thr <= 64.0 * 1.01
    r.reduce_host()
    # lengths 1 and 2 are never removed; length 3 keeps activity >= threshold
    assert r.count(1) == 1 and r.count(2) == 1
    kept = [r.clause_id(3, i) for i in range(r.count(3))]
    assert kept == [0, 6, 7]
    assert r.get_clause(3, 0) == [0, 2, 4]
    assert r.db_clauses() == len(kept) + 2


def test_activity_rescale():
    r = Rig(4, 1, decay=1e-10)
    r.add_clause([0, 2, 4]); r.add_clause([0, 2, 6]); r.add_clause([2, 4, 6])
    r.drain()
    # the increment passed 1e19 on the second clause and everything was rescaled (Clauses.cu:284-291)
    assert r.activity(3, 2) < 1e19 and r.activity(3, 0) < r.activity(3, 1) < r.activity(3, 2)


def _params_tuple(p):
    return (p.startVals, p.lastMask, p.allAggBits, p.usedAggBits, p.updStart, p.updCount, p.nGroups,
            tuple(p.groupAggBit[: p.nGroups]), tuple(p.groupSlotMask[: p.nGroups]))


def test_split_collect_equals_single_pass_collect():
    """Sharer::collectBatch collects large batches in two steps (sizes, then concurrent copies into
    disjoint ranges): same deltas, same run parameters, same assignment ids as the one-pass collect"""
    rng = np.random.default_rng(5)
    nvars, nsolvers = 500, 7
    a, b = Rig(nvars, nsolvers), Rig(nvars, nsolvers)
    for rnd in range(4):
        for s in range(nsolvers):
            for _ in range(int(rng.integers(0, 9))):
                for v in rng.choice(nvars, size=int(rng.integers(0, 60)), replace=False):
                    val = int(rng.integers(0, 3))
                    a.set(s, int(v), val)
                    b.set(s, int(v), val)
                assert a.send(s) == b.send(s)
        ua, pa = a.collect()
        ub, pb = b.collect_split()
        assert np.array_equal(ua, ub)
        assert [_params_tuple(p) for p in pa] == [_params_tuple(p) for p in pb]
        assert [a.ids(s) for s in range(nsolvers)] == [b.ids(s) for s in range(nsolvers)]


def test_take_collect_equals_single_pass_collect():
    """the direct pipeline swaps every solver's delta buffer out (SolverAssigs::takeUpdatesLocked, two
    buffers in rotation) instead of copying it: same deltas, run parameters and assignment ids as the
    one-pass collect, run after run (the rotation re-uses a buffer every second run)"""
    rng = np.random.default_rng(15)
    nvars, nsolvers = 400, 5
    a, b = Rig(nvars, nsolvers), Rig(nvars, nsolvers)
    for rnd in range(7):
        for s in range(nsolvers):
            for _ in range(int(rng.integers(0, 9))):
                for v in rng.choice(nvars, size=int(rng.integers(0, 80)), replace=False):
                    val = int(rng.integers(0, 3))
                    a.set(s, int(v), val)
                    b.set(s, int(v), val)
                assert a.send(s) == b.send(s)
        ua, pa = a.collect()
        ub, pb = b.collect_take()
        assert np.array_equal(ua, ub)
        assert [_params_tuple(p) for p in pa] == [_params_tuple(p) for p in pb]
        assert [a.ids(s) for s in range(nsolvers)] == [b.ids(s) for s in range(nsolvers)]


@pytest.mark.parametrize("parts,share_bound", [(1, 0), (3, 0), (2, 40), (4, 40)])
def test_view_hand_over_equals_host_hand_over(parts, share_bound):
    """the direct pipeline's zero-copy hand-over (Reported::handOverViews: every solver's batch views the
    ids / positions / literal stream the GPU wrote, one slice per device) hands over exactly what the
    host path does, including the re-report rules across runs and the empty progress-marker batches"""
    rng = np.random.default_rng(16)
    nvars, nsolvers = 300, 6
    a, b = Rig(nvars, nsolvers), Rig(nvars, nsolvers)
    counts = {}
    for i in range(3000):
        n = int(rng.integers(1, 7))
        lits = [2 * int(v) + int(rng.integers(0, 2)) for v in rng.choice(nvars, size=n, replace=False)]
        assert a.add_clause(lits) == b.add_clause(lits)
        counts[n] = counts.get(n, 0) + 1
    a.drain(); b.drain()
    for rnd in range(4):
        for r in (a, b):
            for s in range(nsolvers):
                if s != 3:  # one solver without any assignment
                    r.set(s, rnd, 0)
                    r.send(s)
            r.collect()
        hits = set()
        for s in (0, 1, 2, 4, 5):  # solver 3 has no hit; solver 5 few
            for _ in range(3 if s == 5 else 300):
                n = int(rng.integers(1, 7))
                hits.add((s, n, int(rng.integers(0, min(counts[n], 40)))))  # small range: re-reports across runs
        hits = np.array(sorted(hits), dtype=np.int64)
        rec = np.zeros(len(hits), dtype=HIT)
        rec["solver"], rec["len"], rec["idx"] = hits[:, 0], hits[:, 1], hits[:, 2]
        rec["mask"] = rng.integers(1, 1 << 31, size=len(hits))
        a.hand_over(rec[rng.permutation(len(rec))])
        # share_bound: the slices are the devices' shares of every length (several devices behind one front-end):
        # the batch merges them back into the single-device order
        b.hand_over_views(rec[rng.permutation(len(rec))], parts, share_bound)
        for s in range(nsolvers):
            pa, pb = a.pop_all(s), b.pop_all(s)
            assert pa == pb
            assert a.last_all_reported(s) == b.last_all_reported(s)
            assert [a.solver_stat(s, k) for k in range(3)] == [b.solver_stat(s, k) for k in range(3)] if hasattr(a, "solver_stat") else True


def test_sorted_hand_over_equals_host_hand_over():
    """a hit list sorted and resolved 'by the GPU' (ids + literal stream, Reported::handOverSorted: every
    solver's batch is one slice found by binary search) hands over exactly what the host path does"""
    rng = np.random.default_rng(6)
    nvars, nsolvers = 300, 6
    a, b = Rig(nvars, nsolvers), Rig(nvars, nsolvers)
    counts = {}
    for i in range(8000):  # several 128-clause tiles per length
        n = int(rng.integers(1, 7))
        lits = [2 * int(v) + int(rng.integers(0, 2)) for v in rng.choice(nvars, size=n, replace=False)]
        assert a.add_clause(lits) == b.add_clause(lits)
        counts[n] = counts.get(n, 0) + 1
    a.drain(); b.drain()
    for rnd in range(3):
        for r in (a, b):
            for s in range(nsolvers):
                if s != 3:  # one solver without any assignment
                    r.set(s, rnd, 0)
                    r.send(s)
            r.collect()
        hits = []
        for s in (0, 1, 2, 4, 5):  # solver 3 has no hit; solver 5 few
            for _ in range(5 if s == 5 else 4000):
                n = int(rng.integers(1, 7))
                hits.append((int(rng.integers(1, 1 << 31)), s, n, int(rng.integers(0, counts[n]))))
        hits = np.array(sorted(set((s, n, i) for _, s, n, i in hits)), dtype=np.int64)
        rec = np.zeros(len(hits), dtype=HIT)
        rec["solver"], rec["len"], rec["idx"] = hits[:, 0], hits[:, 1], hits[:, 2]
        rec["mask"] = rng.integers(1, 1 << 31, size=len(hits))
        assert len(rec) >= 8192  # the parallel host path
        a.hand_over(rec[rng.permutation(len(rec))])
        b.hand_over_sorted(rec[rng.permutation(len(rec))])
        for s in range(nsolvers):
            pa, pb = a.pop_all(s), b.pop_all(s)
            assert pa == pb
            assert (len(pa) > 0) == (s != 3)
            assert a.last_all_reported(s) == b.last_all_reported(s)


def test_clause_mirror_round_trip_across_tiles():
    """the tiled, lane-interleaved host mirror (clause_db.h wordPos / common.h tileSlot) gives every
    clause back, before and after the first-literal sort of a bulk load"""
    rng = np.random.default_rng(8)
    r = Rig(400, 1)
    want = {}
    offsets, flat = [0], []
    for i in range(9000):  # > 1024 clauses per length: the bulk load is sorted by first literal
        n = int(rng.integers(1, 6))
        lits = [2 * int(v) + int(rng.integers(0, 2)) for v in rng.choice(400, size=n, replace=False)]
        flat += lits
        offsets.append(len(flat))
    first = r.add_clauses_bulk(offsets, flat)
    r.drain()
    for i in range(9000):
        want[first + i] = flat[offsets[i]:offsets[i + 1]]
    seen = 0
    for n in range(1, 6):
        prev_first = -1
        for i in range(r.count(n)):
            lits = r.get_clause(n, i)
            assert lits == want[r.clause_id(n, i)]
            assert lits[0] >= prev_first  # bulk loads are ordered by first literal
            prev_first = lits[0]
            seen += 1
    assert seen == 9000


class _Front:
    """the thin facade logic of Sharer (trySet all-or-nothing, buffered unsets: sharer.cu:148-193) over
    the host rig, so that the rig can be fed the same API traffic as the snapshot model"""

    def __init__(self, rig, nsolvers):
        self.r, self.to_unset = rig, [[] for _ in range(nsolvers)]

    def _flush(self, s):
        for l in self.to_unset[s]:
            self.r.set(s, l >> 1, UNDEF)
        self.to_unset[s] = []

    def trySetSolverValues(self, s, lits):
        if not self.r.available(s):
            return False
        self._flush(s)
        for l in lits:
            self.r.set(s, l >> 1, FALSE if l & 1 else TRUE)
        return True

    def unsetSolverValues(self, s, lits):
        if self.r.available(s):
            self._flush(s)
            for l in lits:
                self.r.set(s, l >> 1, UNDEF)
        else:
            self.to_unset[s].extend(lits)

    def trySendAssignment(self, s):
        return self.r.send(s) if self.r.available(s) else -1


def _two_level_check(dev, params, clauses):
    """numpy statement of what k_filter + k_exact compute FROM THE TABLES (aggregate filter, then the
    exact recurrence on the surviving solvers) -- not the oracle's algorithm"""
    M = 0xFFFFFFFF
    agg_start = 0
    for p in params:
        agg_start |= p.usedAggBits
    out = []
    for cid, lits in clauses:
        a, o = agg_start, 0
        for l in lits:
            v = l >> 1
            f = int(dev.can_true[v]) if l & 1 else int(dev.can_false[v])
            u = int(dev.can_undef[v])
            o = (a & u) | (o & f)
            a &= f
        surv = a | o
        for s, p in enumerate(params):
            if not (surv & p.allAggBits) or p.startVals == 0:
                continue
            a, o = p.startVals, 0
            for l in lits:
                v = l >> 1
                d, t = int(dev.def_[s, v]), int(dev.tru[s, v])
                f = d & (t if l & 1 else ~t) & M
                u = ~d & M
                o = (a & u) | (o & f)
                a &= f
            if a | o:
                out.append((cid, s, a | o))
    return sorted(out)


@pytest.mark.parametrize("nsolvers", [1, 3, 7, 32])
def test_host_logic_plus_table_semantics_match_snapshot_model(nsolvers):
    """CPU-only end to end: random API traffic -> the product's slot machine (C++) -> per-run deltas and
    run parameters -> numpy tables (apply, deferred collapse) -> two-level check from the tables, against
    the snapshot model (every frozen slot is a full assignment, checked by the oracle)"""
    from oracle_lib import SharerModel
    rng = np.random.default_rng(70 + nsolvers)
    nvars = 40
    rig = Rig(nvars, nsolvers)
    front, model, dev = _Front(rig, nsolvers), SharerModel(nvars, nsolvers), DeviceModel(nvars, nsolvers)
    clauses, prev = [], None
    total = 0
    for rnd in range(8):
        for _ in range(int(rng.integers(5, 40))):
            n = int(rng.integers(1, 5))
            lits = [2 * int(rng.integers(0, nvars)) + int(rng.integers(0, 2)) for _ in range(n)]  # duplicates allowed
            clauses.append((model.addClause(lits), lits))
        for s in range(nsolvers):
            for _ in range(int(rng.integers(0, 40 if rnd == 3 else 6))):  # round 3 runs out of slots
                vs = rng.choice(nvars, size=int(rng.integers(0, 20)), replace=False)
                x = rng.random(len(vs))
                unset = [2 * int(v) for v, xx in zip(vs, x) if xx < 0.25]
                sets = [2 * int(v) + int(xx < 0.7) for v, xx in zip(vs, x) if xx >= 0.25]
                front.unsetSolverValues(s, unset); model.unsetSolverValues(s, unset)
                assert front.trySetSolverValues(s, sets) == model.trySetSolverValues(s, sets)
                assert front.trySendAssignment(s) == model.trySendAssignment(s)
        upd, params = rig.collect()
        if prev is not None:
            dev.collapse(*prev)  # the collapse of the previous batch is deferred to the next run
        dev.apply(upd, params)
        prev = (upd, params)
        want = model.run()
        got = _two_level_check(dev, params, clauses)
        assert got == [(int(h["clause_id"]), int(h["solver_id"]), int(h["mask"])) for h in want]
        total += len(got)
    assert total > 0


def _rig_with_acts(acts, decay=1.0, length=3):
    """one clause of `length` literals per entry of acts, bumped to that activity (decay 1: every clause
    starts at 1) -- the fixture of the reference's ClauseActivityLbdTest.cu:109-131"""
    r = Rig(4 * len(acts) + 8, 1, decay=decay)
    lit = 0
    for k, a in enumerate(acts):
        r.add_clause([2 * (lit + j) for j in range(length)])
        lit += length
        r.drain()
        for _ in range(a - 1):
            r.bump(length, k)
    return r


@pytest.mark.parametrize("acts,n,lo,hi", [
    ([1, 2, 3], 2, 2.0, 2.1), ([1, 2, 3], 0, -0.1, 0.1), ([1, 2, 3], 3, 3.0, 3.1),  # testApproxNthLargestAct :164-172
    ([1, 1, 1, 3, 3, 100], 3, 1.0, 1.1),                                            # ...OneBig :174-183
    ([1, 1, 1, 1, 10], 2, 1.0, 1.1),                                                # ...OneVal :185-193
])
def test_approx_nth_activity_like_reference(acts, n, lo, hi):
    """the reference's own expectations for approxNthAct (Clauses.cu:492-525): the n-th smallest activity,
    rounded up to its log-scale bucket"""
    r = _rig_with_acts(acts)
    assert [r.activity(3, k) for k in range(len(acts))] == [float(a) for a in acts]
    assert lo < r.approx_nth_act(n) <= hi


def test_activity_only_threshold_with_huge_differences():
    """ClauseActivityLbdTest.cu:231-254 testActOnlyHugeDifferences: two early clauses (activities 2 and 4
    at decay 0.5), ~56 decays, one late clause -- the removal threshold still separates the two early
    ones (the log-scale buckets cope with 17 orders of magnitude)"""
    import math
    decay = 0.5
    r = Rig(4000, 1, decay=decay)
    r.add_clause([0, 2, 4]); r.drain()                    # activity 2
    r.add_clause([6, 8, 10]); r.drain(); r.bump(3, 1)     # starts at the increment (4), bumped once: 8
    a0, a1 = r.activity(3, 0), r.activity(3, 1)
    assert a0 == 2.0 and a1 == 8.0
    c = int(math.log(1e19 / 100) / math.log(1 / decay))
    for i in range(c):                                    # every added clause decays once (Clauses.cu:334)
        r.add_clause([2 * (20 + i)])                      # unit clauses: never removed, not in the way
    r.add_clause([12, 14, 16]); r.drain()
    big = r.activity(3, 2)
    assert big > 1e15 and r.activity(3, 0) == 2.0         # no rescale happened
    thr = r.approx_nth_act(3 // 2 + c // 2)               # the unit clauses rank above the two early ones
    assert thr > a1
    thr = r.approx_nth_act(1)
    assert a0 < thr <= a0 * 1.1
    thr = r.approx_nth_act(2)
    assert a1 < thr <= a1 * 1.1


def test_record_buckets_scale_with_the_devices_own_share():
    """Multi-GPU: a device checks a contiguous share of the tiles of every length.  The position that picks a hit's
    record bucket (LenDir::ascStart + index, kernels.cu: appendRec) must run over THAT share -- 0 .. localClauses-1 in
    the canonical order (length ascending, index ascending) -- or a device would use 1/N of its buckets."""
    from hostshim_lib import shard_directory
    lens, counts = [2, 3, 5, 9], [1000, 700, 385, 130]
    tile = 128
    for world in (1, 2, 3, 8):
        seen_total = 0
        for rank in range(world):
            rows, local = shard_directory(rank, world, lens, counts)
            assert [r[0] for r in rows] == sorted(lens, reverse=True)  # longest first
            pos = []
            for ln, count, first_tile, local_tiles, asc in sorted(rows):  # ascending length = canonical order
                lo, hi = first_tile * tile, min(count, (first_tile + local_tiles) * tile)
                pos += [asc + idx for idx in range(lo, hi)]
            assert pos == list(range(local)), (world, rank)
            seen_total += local
            if local >= 256:  # every bucket of the 256 is in use
                assert {p * 256 // local for p in pos} == set(range(256))
        assert seen_total == sum(counts)
