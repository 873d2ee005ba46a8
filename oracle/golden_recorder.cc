/*
 * golden_recorder.cc -- TEST INFRASTRUCTURE (BASELINE.json configs[0]).  Runs the reference's own
 * CPU-only CDCL solver (glucose-syrup/simp, compiled in place from /root/reference by
 * oracle/Makefile into oracle/_ref/) on random 3-SAT n=300 m=1278 and records, through the three
 * virtual GPU hooks (glucose-syrup/core/Solver.h:265-268), exactly the calls a GPU-helped solver
 * thread would issue against GpuClauseSharer.h: learned clauses (addClause) and trail snapshots
 * (unsetSolverValues / trySetSolverValues / trySendAssignment), following
 * glucose-syrup/gpu/GpuHelpedSolver.cc:94-157.  The event log is the golden INPUT of the checker;
 * tests/golden/make_config1.py turns it into a fixture with the oracle's expected hits.
 *
 * Event log, one event per line:  c <lits>  learned clause | u <lits> unset | s <lits> set |
 *                                 a  send assignment | r  two gpuRun() calls (every 8 sends)
 */
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "simp/SimpSolver.h"
#include "core/Finisher.h"
#include "gpuShareLib/Utils.h"
#include "utils/Utils.h"

using namespace Glucose;

class RecordingSolver : public SimpSolver {
    FILE *out;
    int trailCopiedUntil = 0;
    int sendsSinceRun = 0;
    long maxClauses;

public:
    long clauses = 0, sends = 0;
    RecordingSolver(Finisher &f, const GpuShare::Logger &l, FILE *o, long maxCl) : SimpSolver(0, f, l), out(o), maxClauses(maxCl) {}

    void emit(char tag, const Lit *lits, int n) {
        fprintf(out, "%c", tag);
        for (int i = 0; i < n; i++) fprintf(out, " %d", toInt(lits[i]));
        fprintf(out, "\n");
    }
    // GpuHelpedSolver.cc:111-121
    void unsetFromTrailForGpu(int level) {
        if (level < decisionLevel() && trailCopiedUntil > trail_lim[level]) {
            emit('u', &trail[trail_lim[level]], trailCopiedUntil - trail_lim[level]);
            trailCopiedUntil = trail_lim[level];
        }
    }
    // GpuHelpedSolver.cc:123-157 (slots never run out: a run is requested every 8 sends)
    bool tryCopyTrailForGpu(int level) override {
        if (clauses >= maxClauses) return false;
        int max;
        if (level < decisionLevel()) {
            unsetFromTrailForGpu(level);
            max = trail_lim[level];
        } else {
            max = trail.size();
        }
        if (trailCopiedUntil < max) {
            emit('s', &trail[trailCopiedUntil], max - trailCopiedUntil);
            trailCopiedUntil = max;
        }
        fprintf(out, "a\n");
        sends++;
        if (++sendsSinceRun == 8) {
            fprintf(out, "r\n");
            sendsSinceRun = 0;
        }
        return true;
    }
    // GpuHelpedSolver.cc:94-101
    void sendClauseToGpu(vec<Lit> &lits, int lbd) override {
        if (clauses >= maxClauses) return;
        emit('c', &lits[0], lits.size());
        clauses++;
    }
    // GpuHelpedSolver.cc:159-162
    void cancelUntil(int level) override {
        if (clauses < maxClauses) unsetFromTrailForGpu(level);
        SimpSolver::cancelUntil(level);
    }
    void finish() { fprintf(out, "r\n"); }
};

int main(int argc, char **argv) {
    int n = argc > 1 ? atoi(argv[1]) : 300, m = argc > 2 ? atoi(argv[2]) : 1278;
    long maxClauses = argc > 3 ? atol(argv[3]) : 1500;
    const char *logPath = argc > 4 ? argv[4] : "config1_events.txt";
    const char *cnfPath = argc > 5 ? argv[5] : "config1.cnf";
    double seed = 91648253; // same LCG family as the solver's own (utils/Utils.h:29-43)
    FILE *out = fopen(logPath, "w"), *cnf = fopen(cnfPath, "w");
    Finisher finisher;
    GpuShare::Logger logger{0, GpuShare::directPrint};
    RecordingSolver S(finisher, logger, out, maxClauses);
    S.parsing = 1;
    S.use_simplification = false;
    for (int i = 0; i < n; i++) S.newVar();
    fprintf(cnf, "p cnf %d %d\n", n, m);
    for (int c = 0; c < m; c++) {
        vec<Lit> cl;
        int vs[3];
        for (int k = 0; k < 3; k++) {
            bool fresh;
            do {
                vs[k] = irand(seed, n);
                fresh = true;
                for (int j = 0; j < k; j++) fresh = fresh && vs[j] != vs[k];
            } while (!fresh);
            bool neg = irand(seed, 2);
            cl.push(mkLit(vs[k], neg));
            fprintf(cnf, "%d ", neg ? -(vs[k] + 1) : vs[k] + 1);
        }
        fprintf(cnf, "0\n");
        S.addClause_(cl);
    }
    fclose(cnf);
    S.parsing = 0;
    S.setConfBudget(20000);
    vec<Lit> dummy;
    lbool ret = S.solveLimited(dummy);
    S.finish();
    fclose(out);
    printf("result %s, %ld learned clauses recorded, %ld assignments, %ld conflicts\n",
           ret == l_True ? "SAT" : (ret == l_False ? "UNSAT" : "UNKNOWN"), S.clauses, S.sends, (long)S.conflicts);
    return 0;
}
