/*
 * gss_synth.h -- deterministic synthetic inputs of the shapes BASELINE.json names.
 * Exported by tests/synthlib/libgss_synth.so (host only, plain g++) for tests and bench.py; not part
 * of the product library nor of the GpuClauseSharer.h surface.  RNG: SplitMix64 with the seeds given by the caller.
 */
#ifndef GPUSHARE_B200_SYNTH_H
#define GPUSHARE_B200_SYNTH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* "Luby-like" clause length mix: len_i = min(max_len, 1 + luby(2, i)) with Glucose's luby
 * (glucose-syrup/core/Solver.cc:1967-1979) -> {2,3,5,9,17,30} at max_len 30, mean 4.41.
 * Returns the total number of literals of clauses [0, nclauses). */
int64_t gss_synth_total_lits(int64_t nclauses, int max_len);

/* planted assignment sigma[v] in {0 true, 1 false} */
void gss_synth_sigma(int nvars, uint64_t seed, uint8_t *sigma);

/* Clauses in CSR form.  var uniform in [0,nvars); the literal is TRUE under sigma with
 * probability p_agree (sigma == NULL: uniform sign).  offsets has nclauses+1 entries. */
void gss_synth_clauses(int64_t nclauses, int nvars, int max_len, const uint8_t *sigma, double p_agree,
                       uint64_t seed, int64_t *offsets, int32_t *lits);

/* One solver's stream of assignments: starts as sigma with each var undefined w.p. p_undef;
 * every step re-draws the status (undefined w.p. p_undef, else sigma's value) of
 * ceil(churn*nvars) randomly chosen variables. */
typedef struct gss_synth_stream gss_synth_stream;
gss_synth_stream *gss_synth_stream_create(int nvars, const uint8_t *sigma, double p_undef, double churn, uint64_t seed);
void gss_synth_stream_destroy(gss_synth_stream *s);
/* current value of every variable: 0 true, 1 false, 2 undef */
const uint8_t *gss_synth_stream_values(gss_synth_stream *s);
/* Advance to the next assignment and return the delta as literals to set (trySetSolverValues)
 * and literals to unset (unsetSolverValues).  The first call emits the whole initial
 * assignment.  Both arrays need room for nvars entries. */
void gss_synth_stream_next(gss_synth_stream *s, int32_t *set_lits, int32_t *n_set, int32_t *unset_lits, int32_t *n_unset);

/* BASELINE configs[3]: a satisfiable CNF instance in DIMACS format for the reference's real solvers:
 * nclauses clauses over nvars variables, lengths from len_weights (len_weights[k] = relative share of
 * clauses with k + 2 literals, n_lens entries), distinct variables per clause, every literal true under
 * the planted assignment with probability p_agree and every clause satisfied by it (one literal is
 * flipped where the draw left none).  Returns 0, or -1 when the file cannot be written. */
int gss_synth_write_cnf(const char *path, int nvars, int64_t nclauses, const double *len_weights, int n_lens,
                        double p_agree, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
