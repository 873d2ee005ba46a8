/*
 * gss_oracle.h -- CPU ORACLE for the GpuShareSat clause-vs-assignment check.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.  The
 * product library (libgpushare_b200.so) never links, loads or calls it.
 *
 * It is a plain-C restatement of the hit semantics of the reference's GPU checker
 * (all citations relative to /root/reference):
 *   - gpuShareLib/GpuRunner.cu:34-57   ReportComputer (allFalse / justOneUndefined recurrence)
 *   - gpuShareLib/GpuRunner.cu:68-89   dCheckOneClauseOneSolver (exact pass on one solver's 32 slots)
 *   - gpuShareLib/GpuRunner.cu:134-177 dFindClauses (aggregate pre-filter, a pure superset filter)
 *   - gpuShareLib/BaseTypes.cuh:65-101 MultiLBool {isDef,isTrue}; :153-159 dSign/dVar
 *   - gpuShareLib/SolverTypes.h:44-60  Lit = 2*var + sign (sign 1 = negated)
 * Parity is PINNED: tests/test_oracle_kat.py reproduces the reference's own known-answer
 * counts 143 (n=15) and 19739 (n=2000) of glucose-syrup/perftest/perfTest.cu:202-204.
 */
#ifndef GSS_ORACLE_H
#define GSS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* value encoding of one variable in one assignment, same as the reference's lbool
 * (gpuShareLib/SolverTypes.h:77-80): 0 true, 1 false, 2 undef */
enum { GSS_TRUE = 0, GSS_FALSE = 1, GSS_UNDEF = 2 };

typedef struct {
    int64_t  clause;  /* index of the clause in the caller's CSR arrays */
    int32_t  solver;
    uint32_t mask;    /* bit p set <=> the clause fires on slot p of that solver */
} gss_oracle_hit;

/* ---- glucose-syrup/utils/Utils.h:29-43 LCG, restated (needed for the KAT inputs) ---- */
double gss_oracle_drand(double *seed);
int    gss_oracle_irand(double *seed, int size);

/* ---- scalar semantics: one clause, one assignment ----
 * returns 1 iff no literal occurrence is true and at most one is undefined
 * (GpuRunner.cu:49-56: the reported set is allFalse | justOneUndefined). */
int gss_oracle_clause_fires(const int32_t *lits, int n, const uint8_t *vals);

/* ---- bit-parallel over the 32 slots of one solver (GpuRunner.cu:68-89) ----
 * def/tru: per variable 32-bit words; start: slots that take part. */
uint32_t gss_oracle_clause_mask32(const int32_t *lits, int n, const uint32_t *def,
                                  const uint32_t *tru, uint32_t start);

/* ---- whole clause DB against nsolvers x 32 slots ----
 * Clauses in CSR form (offsets[nclauses+1], lits[]).  def/tru are [nsolvers][nvars].
 * start[s] = slots of solver s that take part.  Hits are written in (clause, solver)
 * order; returns the total number of hits (may exceed cap; only cap are stored).
 * use_filter != 0 follows the reference's two-level structure (one aggregate bit per
 * solver first, GpuRunner.cu:148-172, then the exact per-solver pass); the result is
 * identical either way.  nthreads <= 1 runs inline. */
int64_t gss_oracle_check_db(const int64_t *offsets, const int32_t *lits, int64_t nclauses,
                            int nsolvers, int64_t nvars, const uint32_t *def, const uint32_t *tru,
                            const uint32_t *start, gss_oracle_hit *out, int64_t cap,
                            int use_filter, int nthreads);
/* bench.py's CPU arm: the same check with the two-level structure on, the per-variable aggregates
 * built by all `nthreads` threads instead of serially (identical hits). */
int64_t gss_oracle_check_db_bench(const int64_t *offsets, const int32_t *lits, int64_t nclauses,
                                  int nsolvers, int64_t nvars, const uint32_t *def, const uint32_t *tru,
                                  const uint32_t *start, gss_oracle_hit *out, int64_t cap, int nthreads);

/* ---- the reference's perf-test known-answer inputs (perftest/perfTest.cu:70-107) ---- */
/* clauses: seed 0.4; size = irand(seed,minLen,maxLen); per literal sign drawn BEFORE var
 * (g++ evaluates mkLit(irand(varCount), irand(2)) right to left). offsets has nclauses+1
 * entries, lits must hold nclauses*(maxLen-1) entries.  Returns literal count. */
int64_t gss_oracle_kat_clauses(int64_t nclauses, int minLen, int maxLen, int nvars, double *seed,
                               int64_t *offsets, int32_t *lits);
/* one assignment: per var p = irand(seed,3): 0 -> true, 1 -> false, 2 -> undef */
void gss_oracle_kat_assignment(double *seed, int nvars, uint8_t *vals);
/* the whole KAT on the CPU: n sweeps, one assignment each; returns the number of
 * (clause, assignment) hits; *all_false gets the number with no undefined literal. */
int64_t gss_oracle_kat_run(int64_t nclauses, int minLen, int maxLen, int nvars, int n, int scalar,
                           int64_t *all_false);

#ifdef __cplusplus
}
#endif
#endif
