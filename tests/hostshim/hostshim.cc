// hostshim.cc -- TEST-ONLY C hooks over the host classes of the product (assigs.cc, reported.cc,
// clause_db.cc) so that `pytest -m "not gpu"` can exercise the slot state machine, the hand-over
// rules and the reduceDb host logic without a CUDA device.  Not part of libgpushare_b200.so.
#include "../../gpusharesat_b200/csrc/assigs.h"
#include "../../gpusharesat_b200/csrc/clause_db.h"
#include "../../gpusharesat_b200/csrc/reported.h"
#include "../../gpusharesat_b200/csrc/stats.h"
#include <algorithm>
#include <chrono>
#include <cstring>

using namespace gss;

// the host-only rig never has activities on a device (kernels.cu is not linked here)
namespace gss {
void scaleActivitiesOnDevice(float *, int64_t, float, cudaStream_t) {}
// (reduce.cu is device code)
struct ClauseDb::PermScratch {};
void ClauseDb::PermScratchDeleter::operator()(PermScratch *p) const { delete p; }
void ClauseDb::initPermScratch() {}
void ClauseDb::releaseSpare(cudaStream_t) {}
bool ClauseDb::permuteOnDevice(cudaStream_t, bool) { return false; }
} // namespace gss

struct HostRig {
    Logger logger;
    HostAssigs assigs;
    ClauseDb db;
    std::vector<std::vector<uint64_t>> stats;
    Reported reported;
    HostBuf<VarUpdate> updates;
    std::vector<SolverRunParams> params;
    std::vector<AssigIds> ids;
    HostRig(int nvars, int nsolvers, double decay) : db(decay, logger, 0), reported(db, stats) {
        logger.verbosity = 0;
        assigs.setVarCount(nvars);
        assigs.growSolvers(nsolvers);
        stats.assign(nsolvers, std::vector<uint64_t>(S_COUNT, 0));
        reported.setSolverCount(nsolvers);
    }
};

extern "C" {

HostRig *hs_create(int nvars, int nsolvers, double decay) { return new HostRig(nvars, nsolvers, decay); }
void hs_destroy(HostRig *r) { delete r; }

void hs_set_host_bumps(HostRig *r, int on) { r->reported.setHostBumps(on != 0); }

// ---- slot state machine ----
int hs_available(HostRig *r, int s) { return r->assigs.solver(s).isAssignmentAvailableLocked() ? 1 : 0; }
void hs_set_var(HostRig *r, int s, int var, int val) { r->assigs.solver(s).setVarLocked(var, (uint8_t)val); }
int64_t hs_send(HostRig *r, int s) {
    int64_t id = r->assigs.solver(s).assignmentDoneLocked();
    r->reported.assigWasSent(s, id);
    return id;
}
// one run's collection for every solver; returns the number of updates
int hs_collect(HostRig *r, int fullRebuild) {
    int n = r->assigs.solverCount();
    r->updates.clear();
    r->params.assign(n, SolverRunParams{});
    r->ids.assign(n, AssigIds{});
    for (int s = 0; s < n; s++) r->assigs.solver(s).collectLocked(r->updates, r->params[s], r->ids[s], fullRebuild != 0);
    return (int)r->updates.size();
}
// the same collection the way the engine does it for large batches (Sharer::collectBatch): one pass
// for the sizes, then every solver copies into its own range -- here on a worker pool
int hs_collect_split(HostRig *r) {
    int n = r->assigs.solverCount();
    r->updates.clear();
    r->params.assign(n, SolverRunParams{});
    r->ids.assign(n, AssigIds{});
    std::vector<size_t> offset(n + 1, 0);
    size_t total = 0;
    for (int s = 0; s < n; s++) {
        offset[s] = total;
        total += r->assigs.solver(s).pendingUpdatesLocked();
    }
    VarUpdate *base = r->updates.append(total);
    LazyPool pool;
    pool.get().parallelFor(n, [&](int s) {
        r->assigs.solver(s).collectIntoLocked(base + offset[s], (int32_t)offset[s], r->params[s], r->ids[s]);
    });
    return (int)total;
}
// the direct pipeline's collection (Sharer::startRunDirect): every solver's delta buffer is swapped out
// (SolverAssigs::takeUpdatesLocked) and read where it lies; here the taken buffers are concatenated so
// that the result can be compared with hs_collect
int hs_collect_take(HostRig *r) {
    int n = r->assigs.solverCount();
    r->updates.clear();
    r->params.assign(n, SolverRunParams{});
    r->ids.assign(n, AssigIds{});
    int64_t total = 0;
    for (int s = 0; s < n; s++) {
        const VarUpdate *ptr = nullptr;
        bool pinned = false;
        r->assigs.solver(s).takeUpdatesLocked(ptr, (int32_t)total, r->params[s], r->ids[s], &pinned);
        int cnt = r->params[s].updCount;
        if (cnt) memcpy(r->updates.append((size_t)cnt), ptr, (size_t)cnt * sizeof(VarUpdate));
        total += cnt;
    }
    return (int)total;
}
void hs_get_params(HostRig *r, int s, SolverRunParams *out) { *out = r->params[s]; }
void hs_get_updates(HostRig *r, VarUpdate *out) { memcpy(out, r->updates.data(), r->updates.size() * sizeof(VarUpdate)); }
void hs_get_ids(HostRig *r, int s, int64_t *start, int *count) { *start = r->ids[s].start; *count = r->ids[s].count; }

// ---- clause database (host half) ----
int64_t hs_add_clause(HostRig *r, const int *lits, int n) { return r->db.addClause(lits, n); }
void hs_drain(HostRig *r) { r->db.drainPending(); }
int hs_count(HostRig *r, int len) { return r->db.count(len); }
int64_t hs_clause_id(HostRig *r, int len, int idx) { return r->db.clauseId(len, idx); }
float hs_activity(HostRig *r, int len, int idx) { return r->db.activity(len, idx); }
void hs_bump(HostRig *r, int len, int idx) { r->db.bumpActivity(len, idx); }
float hs_approx_nth_act(HostRig *r, int64_t n) { return r->db.approxNthAct(n); }
void hs_reduce_host(HostRig *r) { r->db.reduceHost(); }
int64_t hs_db_clauses(HostRig *r) { return r->db.stats().clauses; }
int64_t hs_db_length_sum(HostRig *r) { return r->db.stats().lengthSum; }
int hs_get_clause(HostRig *r, int len, int idx, int *out) {
    std::vector<int> l;
    int64_t id;
    r->db.getClause(len, idx, l, id);
    memcpy(out, l.data(), l.size() * sizeof(int));
    return (int)l.size();
}

// ---- hand-over ----
void hs_clause_was_added(HostRig *r, int s, int64_t id) { r->reported.clauseWasAdded(s, id); }
// hits: (mask, solver, len, idx) quadruples; uses the ids of the last hs_collect
void hs_fill(HostRig *r, const HitRecord *hits, int n) { r->reported.fill(r->ids, hits, (size_t)n); }
// the whole post-run host path (sort, bumps, batches; parallel for large lists); returns microseconds
double hs_hand_over(HostRig *r, const HitRecord *hits, int n) {
    std::vector<HitRecord> v(hits, hits + n);
    auto t0 = std::chrono::steady_clock::now();
    r->reported.handOver(v, r->ids, r->assigs.solverCount());
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
}
// the hand-over of a hit list the GPU has already sorted and resolved (Reported::handOverSorted): the
// records, ids and the literal stream are built here from the host mirror, in (solver, len, idx) order
double hs_hand_over_sorted(HostRig *r, const HitRecord *hits, int n) {
    std::vector<HitRecord> v(hits, hits + n);
    std::sort(v.begin(), v.end(), [](const HitRecord &a, const HitRecord &b) {
        if (a.solver != b.solver) return a.solver < b.solver;
        if (a.len != b.len) return a.len < b.len;
        return a.idx < b.idx;
    });
    std::vector<SortedHit> recs(v.size());
    std::vector<int> lits;
    for (size_t i = 0; i < v.size(); i++) {
        const int64_t pos = (int64_t)lits.size();
        const int64_t id = r->db.appendClause(v[i].len, v[i].idx, lits);
        recs[i] = SortedHit{v[i].mask, v[i].solver, v[i].len, v[i].idx, id, pos};
    }
    auto t0 = std::chrono::steady_clock::now();
    r->reported.handOverSorted(recs.data(), recs.size(), lits.data(), (int64_t)lits.size(), r->ids, r->assigs.solverCount());
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
}
// the direct pipeline's hand-over (Reported::handOverViews): per solver the result exactly as k_emit
// lays it out in a result buffer -- ids, n + 1 positions, literal stream -- split in `parts` slices
// (one per device in a multi-GPU run); the batches view the buffer, which lives until they let go
// shareBound > 0: the slices are cut the way devices share the work -- part p holds the clauses of EVERY length
// whose index falls into the p-th share [bound*p/parts, bound*(p+1)/parts) -- so the canonical order interleaves
// the parts and the batch has to merge its views (ClauseBatch::pop)
double hs_hand_over_views_shares(HostRig *r, const HitRecord *hits, int n, int parts, int shareBound) {
    std::vector<HitRecord> v(hits, hits + n);
    std::sort(v.begin(), v.end(), [](const HitRecord &a, const HitRecord &b) {
        if (a.solver != b.solver) return a.solver < b.solver;
        if (a.len != b.len) return a.len < b.len;
        return a.idx < b.idx;
    });
    struct Buf {
        std::vector<int64_t> ids;
        std::vector<int32_t> pos, lits;
    };
    const int S = r->assigs.solverCount();
    std::vector<std::vector<ResultView>> views((size_t)S);
    size_t i = 0;
    for (int s = 0; s < S; s++) {
        size_t lo = i;
        while (i < v.size() && v[i].solver == s) i++;
        const size_t cnt = i - lo;
        for (int part = 0; part < parts; part++) {
            const size_t a = shareBound > 0 ? lo : lo + cnt * part / parts, b = shareBound > 0 ? i : lo + cnt * (part + 1) / parts;
            auto buf = std::make_shared<Buf>();
            std::vector<int> tmp;
            for (size_t k = a; k < b; k++) {
                if (shareBound > 0 && std::min(parts - 1, (int)((int64_t)v[k].idx * parts / shareBound)) != part) continue;
                buf->pos.push_back((int32_t)buf->lits.size());
                tmp.clear();
                buf->ids.push_back(r->db.appendClause(v[k].len, v[k].idx, tmp));
                buf->lits.insert(buf->lits.end(), tmp.begin(), tmp.end());
            }
            buf->pos.push_back((int32_t)buf->lits.size());
            ResultView rv;
            rv.ids = buf->ids.data();
            rv.pos = buf->pos.data();
            rv.lits = buf->lits.data();
            rv.n = (int32_t)buf->ids.size();
            rv.owner = buf;
            views[s].push_back(std::move(rv));
        }
    }
    auto t0 = std::chrono::steady_clock::now();
    r->reported.handOverViews(views, r->ids, S);
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
}
double hs_hand_over_views(HostRig *r, const HitRecord *hits, int n, int parts) { return hs_hand_over_views_shares(r, hits, n, parts, 0); }
int64_t hs_add_clauses_bulk(HostRig *r, const int64_t *offsets, const int *lits, int64_t n) { return r->db.addClausesBulk(offsets, lits, n); }
int hs_pop(HostRig *r, int s, int *lits, int *count, int64_t *id) {
    int *l;
    int c;
    int64_t i;
    if (!r->reported.pop(s, l, c, i)) return 0;
    memcpy(lits, l, c * sizeof(int));
    *count = c;
    *id = i;
    return 1;
}
int64_t hs_last_all_reported(HostRig *r, int s) { return r->reported.lastAssigAllReported(s); }
int64_t hs_solver_stat(HostRig *r, int s, int stat) { return (int64_t)r->stats[s][stat]; }


// ---- directory of a device's share (multi-GPU): one row per non-empty length, longest first ----
// counts[i] clauses of length lens[i] are added to a fresh database with shard (rank, world); out = rows of
// {len, count, firstTile, localTiles, ascStart}; returns the number of rows, *localTotal = clauses the device checks
int hs_shard_directory(int rank, int world, const int *lens, const int *counts, int n, int64_t *out, int64_t *localTotal) {
    Logger logger;
    logger.verbosity = 0;
    ClauseDb db(0.999, logger, 0);
    db.setShard(rank, world);
    std::vector<int> lits;
    for (int i = 0; i < n; i++) {
        lits.assign((size_t)lens[i], 2);
        for (int c = 0; c < counts[i]; c++) db.addClause(lits.data(), lens[i]);
    }
    db.drainPending();
    std::vector<LenDir> dir;
    db.buildDirectory(dir);
    int prevEnd = 0;
    for (size_t k = 0; k < dir.size(); k++) {
        out[5 * k] = dir[k].len;
        out[5 * k + 1] = dir[k].count;
        out[5 * k + 2] = dir[k].firstTile;
        out[5 * k + 3] = dir[k].tileEnd - prevEnd;
        out[5 * k + 4] = dir[k].ascStart;
        prevEnd = dir[k].tileEnd;
    }
    *localTotal = db.localClauses();
    return (int)dir.size();
}

} // extern "C"
