/*
 * gss_oracle.c -- CPU ORACLE (test infrastructure only; see gss_oracle.h for the rules).
 * Plain-C restatement of the reference's clause-vs-assignment check.  Every function
 * cites the reference lines it follows (paths relative to /root/reference).
 */
#include "gss_oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* glucose-syrup/utils/Utils.h:29-33 */
double gss_oracle_drand(double *seed) {
    *seed *= 1389796;
    int q = (int)(*seed / 2147483647);
    *seed -= (double)q * 2147483647;
    return *seed / 2147483647;
}

/* glucose-syrup/utils/Utils.h:36-38 */
int gss_oracle_irand(double *seed, int size) { return (int)(gss_oracle_drand(seed) * size); }

/* Scalar statement of gpuShareLib/GpuRunner.cu:49-56 for a single assignment:
 * the clause is reported iff it ends in state allFalse or justOneUndefined. */
int gss_oracle_clause_fires(const int32_t *lits, int n, const uint8_t *vals) {
    int undef = 0;
    for (int i = 0; i < n; i++) {
        int32_t lit = lits[i];
        int v = lit >> 1, neg = lit & 1; /* SolverTypes.h:57-58 */
        uint8_t val = vals[v];
        if (val == GSS_UNDEF) {
            if (++undef > 1) return 0;
        } else {
            int lit_true = neg ? (val == GSS_FALSE) : (val == GSS_TRUE);
            if (lit_true) return 0;
        }
    }
    return 1;
}

/* gpuShareLib/GpuRunner.cu:68-89 (dCheckOneClauseOneSolver) + :49-52 (update) */
uint32_t gss_oracle_clause_mask32(const int32_t *lits, int n, const uint32_t *def,
                                  const uint32_t *tru, uint32_t start) {
    uint32_t all_false = start, just_one_undef = 0;
    for (int i = 0; i < n; i++) {
        int32_t lit = lits[i];
        int v = lit >> 1;
        uint32_t is_false = tru[v];
        if (!(lit & 1)) is_false = ~is_false; /* positive literal is false where the var is false */
        uint32_t d = def[v];
        uint32_t can_be_false = is_false & d, can_be_undef = ~d;
        just_one_undef = (all_false & can_be_undef) | (just_one_undef & can_be_false);
        all_false &= can_be_false;
        if (!(all_false | just_one_undef)) return 0;
    }
    return all_false | just_one_undef;
}

/* ---------------- whole-DB check ---------------- */

typedef struct {
    const int64_t *offsets;
    const int32_t *lits;
    int64_t c0, c1;
    int nsolvers;
    int64_t nvars;
    const uint32_t *def, *tru, *start;
    /* aggregate filter: per var {can_be_true, can_be_false, can_be_undef}, bit groups per solver */
    const uint32_t *agg; /* [nvars][3] or NULL */
    uint32_t agg_start;
    const uint32_t *solver_bits; /* [nsolvers] mask of aggregate bits owned by each solver */
    gss_oracle_hit *out;
    int64_t cap, count;
} job_t;

static void *run_job(void *p) {
    job_t *j = (job_t *)p;
    j->count = 0;
    for (int64_t c = j->c0; c < j->c1; c++) {
        const int32_t *cl = j->lits + j->offsets[c];
        int n = (int)(j->offsets[c + 1] - j->offsets[c]);
        uint32_t surviving = 0xFFFFFFFFu;
        if (j->agg) {
            /* GpuRunner.cu:148-172: the same recurrence over the per-var aggregates */
            uint32_t all_false = j->agg_start, just_one = 0;
            for (int i = 0; i < n && (all_false | just_one); i++) {
                int32_t lit = cl[i];
                const uint32_t *a = j->agg + 3 * (int64_t)(lit >> 1);
                uint32_t f = (lit & 1) ? a[0] : a[1];
                just_one = (all_false & a[2]) | (just_one & f);
                all_false &= f;
            }
            surviving = all_false | just_one;
            if (!surviving) continue;
        }
        for (int s = 0; s < j->nsolvers; s++) {
            if (j->agg && !(surviving & j->solver_bits[s])) continue; /* GpuRunner.cu:116-131 */
            if (!j->start[s]) continue;
            uint32_t m = gss_oracle_clause_mask32(cl, n, j->def + (int64_t)s * j->nvars,
                                                  j->tru + (int64_t)s * j->nvars, j->start[s]);
            if (m) {
                if (j->count < j->cap) {
                    gss_oracle_hit *h = &j->out[j->count];
                    h->clause = c;
                    h->solver = s;
                    h->mask = m;
                }
                j->count++;
            }
        }
    }
    return NULL;
}

/* Build per-var aggregates the way gpuShareLib/Assigs.cu:54-69,402-426 partitions 32
 * aggregate bits among the solvers and Assigs.cu:263-284 spreads a solver's live slots
 * over its bits.  Any grouping gives a superset filter; results do not depend on it. */
static uint32_t *build_aggregates(int nsolvers, int64_t nvars, const uint32_t *def,
                                  const uint32_t *tru, const uint32_t *start,
                                  uint32_t *agg_start, uint32_t *solver_bits) {
    uint32_t *agg = (uint32_t *)calloc((size_t)nvars * 3, sizeof(uint32_t));
    int low = 32 / nsolvers, missing = 32 - low * nsolvers, bit = 0;
    *agg_start = 0;
    for (int s = 0; s < nsolvers; s++) {
        int nbits = low + (s < missing ? 1 : 0);
        int live = __builtin_popcount(start[s]);
        int used = nbits < live ? nbits : live;
        solver_bits[s] = 0;
        for (int b = 0; b < nbits; b++) solver_bits[s] |= 1u << (bit + b);
        if (used > 0) {
            /* slot groups in bit order */
            uint32_t group_mask[32];
            int per = live / used, extra = live - per * used, pos = 0;
            for (int g = 0; g < used; g++) {
                int want = per + (g < extra ? 1 : 0);
                uint32_t m = 0;
                while (want > 0) {
                    if (start[s] & (1u << pos)) { m |= 1u << pos; want--; }
                    pos++;
                }
                group_mask[g] = m;
                *agg_start |= 1u << (bit + g);
            }
            const uint32_t *d = def + (int64_t)s * nvars, *t = tru + (int64_t)s * nvars;
            for (int64_t v = 0; v < nvars; v++) {
                uint32_t bt = t[v] & d[v], bf = ~t[v] & d[v], bu = ~d[v];
                uint32_t *a = agg + 3 * v;
                for (int g = 0; g < used; g++) {
                    uint32_t gb = 1u << (bit + g);
                    if (bt & group_mask[g]) a[0] |= gb;
                    if (bf & group_mask[g]) a[1] |= gb;
                    if (bu & group_mask[g]) a[2] |= gb;
                }
            }
        }
        bit += nbits;
    }
    return agg;
}

/* bench.py's CPU arm only: the same aggregates, built by `nthreads` threads (each owns a range of
 * variables).  In the reference the aggregates are maintained incrementally on the device
 * (dUpdateAssigs, Assigs.cu:100-116), so a serial rebuild per sweep would handicap the CPU arm. */
typedef struct {
    int nsolvers; int64_t nvars, v0, v1;
    const uint32_t *def, *tru, *start;
    uint32_t *agg;
} agg_job_t;

static void *run_agg_job(void *arg) {
    agg_job_t *j = (agg_job_t *)arg;
    int low = 32 / j->nsolvers, missing = 32 - low * j->nsolvers, bit = 0;
    for (int s = 0; s < j->nsolvers; s++) {
        int nbits = low + (s < missing ? 1 : 0);
        int live = __builtin_popcount(j->start[s]);
        int used = nbits < live ? nbits : live;
        if (used > 0) {
            uint32_t group_mask[32];
            int per = live / used, extra = live - per * used, pos = 0;
            for (int g = 0; g < used; g++) {
                int want = per + (g < extra ? 1 : 0);
                uint32_t m = 0;
                while (want > 0) {
                    if (j->start[s] & (1u << pos)) { m |= 1u << pos; want--; }
                    pos++;
                }
                group_mask[g] = m;
            }
            const uint32_t *d = j->def + (int64_t)s * j->nvars, *t = j->tru + (int64_t)s * j->nvars;
            for (int64_t v = j->v0; v < j->v1; v++) {
                uint32_t bt = t[v] & d[v], bf = ~t[v] & d[v], bu = ~d[v];
                uint32_t *a = j->agg + 3 * v;
                for (int g = 0; g < used; g++) {
                    uint32_t gb = 1u << (bit + g);
                    if (bt & group_mask[g]) a[0] |= gb;
                    if (bf & group_mask[g]) a[1] |= gb;
                    if (bu & group_mask[g]) a[2] |= gb;
                }
            }
        }
        bit += nbits;
    }
    return NULL;
}

static uint32_t *build_aggregates_mt(int nsolvers, int64_t nvars, const uint32_t *def, const uint32_t *tru,
                                     const uint32_t *start, uint32_t *agg_start, uint32_t *solver_bits, int nthreads) {
    /* masks and bit ownership: the serial routine on zero variables */
    free(build_aggregates(nsolvers, 0, def, tru, start, agg_start, solver_bits));
    uint32_t *agg = (uint32_t *)calloc((size_t)nvars * 3 + 1, sizeof(uint32_t));
    if (nthreads > nvars) nthreads = nvars > 0 ? (int)nvars : 1;
    agg_job_t *jobs = (agg_job_t *)calloc((size_t)nthreads, sizeof(agg_job_t));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    for (int t = 0; t < nthreads; t++) {
        agg_job_t *j = &jobs[t];
        j->nsolvers = nsolvers; j->nvars = nvars;
        j->v0 = nvars * t / nthreads; j->v1 = nvars * (t + 1) / nthreads;
        j->def = def; j->tru = tru; j->start = start; j->agg = agg;
        pthread_create(&th[t], NULL, run_agg_job, j);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return agg;
}

static int64_t check_db_impl(const int64_t *offsets, const int32_t *lits, int64_t nclauses,
                             int nsolvers, int64_t nvars, const uint32_t *def, const uint32_t *tru,
                             const uint32_t *start, gss_oracle_hit *out, int64_t cap,
                             int use_filter, int nthreads, int parallel_aggregates);

int64_t gss_oracle_check_db(const int64_t *offsets, const int32_t *lits, int64_t nclauses,
                            int nsolvers, int64_t nvars, const uint32_t *def, const uint32_t *tru,
                            const uint32_t *start, gss_oracle_hit *out, int64_t cap,
                            int use_filter, int nthreads) {
    return check_db_impl(offsets, lits, nclauses, nsolvers, nvars, def, tru, start, out, cap, use_filter, nthreads, 0);
}

int64_t gss_oracle_check_db_bench(const int64_t *offsets, const int32_t *lits, int64_t nclauses,
                                  int nsolvers, int64_t nvars, const uint32_t *def, const uint32_t *tru,
                                  const uint32_t *start, gss_oracle_hit *out, int64_t cap, int nthreads) {
    return check_db_impl(offsets, lits, nclauses, nsolvers, nvars, def, tru, start, out, cap, 1, nthreads, 1);
}

static int64_t check_db_impl(const int64_t *offsets, const int32_t *lits, int64_t nclauses,
                             int nsolvers, int64_t nvars, const uint32_t *def, const uint32_t *tru,
                             const uint32_t *start, gss_oracle_hit *out, int64_t cap,
                             int use_filter, int nthreads, int parallel_aggregates) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    uint32_t agg_start = 0, solver_bits[32];
    uint32_t *agg = NULL;
    if (use_filter && nsolvers >= 1 && nsolvers <= 32)
        agg = parallel_aggregates && nthreads > 1
                  ? build_aggregates_mt(nsolvers, nvars, def, tru, start, &agg_start, solver_bits, nthreads)
                  : build_aggregates(nsolvers, nvars, def, tru, start, &agg_start, solver_bits);
    if (nclauses < nthreads) nthreads = nclauses > 0 ? (int)nclauses : 1;

    job_t *jobs = (job_t *)calloc((size_t)nthreads, sizeof(job_t));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    /* each thread writes into a private slice of a scratch buffer, merged in clause order */
    gss_oracle_hit **bufs = (gss_oracle_hit **)calloc((size_t)nthreads, sizeof(*bufs));
    for (int t = 0; t < nthreads; t++) {
        job_t *j = &jobs[t];
        j->offsets = offsets; j->lits = lits;
        j->c0 = nclauses * t / nthreads; j->c1 = nclauses * (t + 1) / nthreads;
        j->nsolvers = nsolvers; j->nvars = nvars;
        j->def = def; j->tru = tru; j->start = start;
        j->agg = agg; j->agg_start = agg_start; j->solver_bits = solver_bits;
        j->cap = cap;
        bufs[t] = nthreads == 1 ? out : (gss_oracle_hit *)malloc((size_t)(cap > 0 ? cap : 1) * sizeof(gss_oracle_hit));
        j->out = bufs[t];
    }
    if (nthreads == 1) {
        run_job(&jobs[0]);
    } else {
        for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, run_job, &jobs[t]);
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    }
    int64_t total = 0;
    for (int t = 0; t < nthreads; t++) {
        if (nthreads > 1) {
            int64_t stored = jobs[t].count < cap ? jobs[t].count : cap;
            int64_t room = cap - total;
            if (room < 0) room = 0;
            if (stored > room) stored = room;
            if (stored > 0) memcpy(out + total, bufs[t], (size_t)stored * sizeof(gss_oracle_hit));
            free(bufs[t]);
        }
        total += jobs[t].count;
    }
    free(bufs); free(th); free(jobs); free(agg);
    return total;
}

/* ---------------- perfTest known-answer inputs ---------------- */

/* glucose-syrup/perftest/perfTest.cu:94-107 (PerfFixture): seed carried by the caller */
int64_t gss_oracle_kat_clauses(int64_t nclauses, int minLen, int maxLen, int nvars, double *seed,
                               int64_t *offsets, int32_t *lits) {
    int64_t pos = 0;
    for (int64_t c = 0; c < nclauses; c++) {
        offsets[c] = pos;
        int size = gss_oracle_irand(seed, maxLen - minLen) + minLen; /* Utils.h:41-43 */
        for (int l = 0; l < size; l++) {
            /* mkLit(irand(seed,nVars), irand(seed,2)): g++ evaluates the arguments right to
             * left, so the sign is drawn first (var-first gives 137, not 143) */
            int sign = gss_oracle_irand(seed, 2);
            int var = gss_oracle_irand(seed, nvars);
            lits[pos++] = var + var + sign;
        }
    }
    offsets[nclauses] = pos;
    return pos;
}

/* glucose-syrup/perftest/perfTest.cu:70-83: p==0 enqueues mkLit(var,false) (var true),
 * p==1 enqueues mkLit(var,true) (var false), p==2 leaves the var unassigned */
void gss_oracle_kat_assignment(double *seed, int nvars, uint8_t *vals) {
    for (int v = 0; v < nvars; v++) {
        int p = gss_oracle_irand(seed, 3);
        vals[v] = p == 0 ? GSS_TRUE : (p == 1 ? GSS_FALSE : GSS_UNDEF);
    }
}

/* glucose-syrup/perftest/perfTest.cu:153-207 (testPerf) on the CPU */
int64_t gss_oracle_kat_run(int64_t nclauses, int minLen, int maxLen, int nvars, int n, int scalar,
                           int64_t *all_false) {
    int64_t *offsets = (int64_t *)malloc((size_t)(nclauses + 1) * sizeof(int64_t));
    int32_t *lits = (int32_t *)malloc((size_t)nclauses * (size_t)maxLen * sizeof(int32_t));
    double cseed = 0.4, aseed = 0.6;
    gss_oracle_kat_clauses(nclauses, minLen, maxLen, nvars, &cseed, offsets, lits);
    uint8_t *vals = (uint8_t *)malloc((size_t)nvars);
    int64_t hits = 0, af = 0;
    if (scalar) {
        for (int it = 0; it < n; it++) {
            gss_oracle_kat_assignment(&aseed, nvars, vals);
            for (int64_t c = 0; c < nclauses; c++) {
                const int32_t *cl = lits + offsets[c];
                int len = (int)(offsets[c + 1] - offsets[c]);
                if (gss_oracle_clause_fires(cl, len, vals)) {
                    hits++;
                    int u = 0;
                    for (int i = 0; i < len; i++) u += vals[cl[i] >> 1] == GSS_UNDEF;
                    af += u == 0;
                }
            }
        }
    } else {
        /* 32 sweeps at a time, one per slot */
        uint32_t *def = (uint32_t *)malloc((size_t)nvars * 4), *tru = (uint32_t *)malloc((size_t)nvars * 4);
        for (int it0 = 0; it0 < n; it0 += 32) {
            int k = n - it0 < 32 ? n - it0 : 32;
            memset(def, 0, (size_t)nvars * 4); memset(tru, 0, (size_t)nvars * 4);
            for (int p = 0; p < k; p++) {
                gss_oracle_kat_assignment(&aseed, nvars, vals);
                for (int v = 0; v < nvars; v++) {
                    if (vals[v] != GSS_UNDEF) def[v] |= 1u << p;
                    if (vals[v] == GSS_TRUE) tru[v] |= 1u << p;
                }
            }
            uint32_t start = k == 32 ? 0xFFFFFFFFu : ((1u << k) - 1);
            for (int64_t c = 0; c < nclauses; c++) {
                const int32_t *cl = lits + offsets[c];
                int len = (int)(offsets[c + 1] - offsets[c]);
                uint32_t m = gss_oracle_clause_mask32(cl, len, def, tru, start);
                if (m) {
                    hits += __builtin_popcount(m);
                    uint32_t nou = m;
                    for (int i = 0; i < len; i++) nou &= def[cl[i] >> 1];
                    af += __builtin_popcount(nou);
                }
            }
        }
        free(def); free(tru);
    }
    free(vals); free(lits); free(offsets);
    if (all_false) *all_false = af;
    return hits;
}
