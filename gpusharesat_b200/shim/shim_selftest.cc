// shim_selftest.cc -- drives the library the way a consumer of the reference does: through the
// reference's own header and factory only (no gss_* call).  Scenario = the reference's
// testClausesAssigsReported (glucose-syrup/test/GpuSolverTest.cu:343-391).
#include <cstdint>
#include <cstdio>
#include <functional>
#include <memory>
#include <string>

#include "gpuShareLib/GpuClauseSharer.h"

using namespace GpuShare;

static int lit(int var, bool neg = false) { return 2 * var + (neg ? 1 : 0); }

static int popCount(GpuClauseSharer &sh, int solver) {
    int n = 0, count, *lits;
    long id;
    while (sh.popReportedClause(solver, lits, count, id)) n++;
    return n;
}

#define CHECK(c) do { if (!(c)) { printf("FAILED: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

int main() {
    GpuClauseSharerOptions opts;
    opts.gpuBlockCountGuideline = 3;
    opts.gpuThreadsPerBlockGuideline = 32;
    opts.verbosity = 0;
    std::unique_ptr<GpuClauseSharer> sh(makeGpuClauseSharerPtr(opts, [](const std::string &) {}));
    sh->setVarCount(3);
    sh->setCpuSolverCount(3);
    for (int v = 0; v < 3; v++) { int l = lit(v); CHECK(sh->addClause(-1, &l, 1) == v); }
    int a0[] = {lit(0, true), lit(1), lit(2)};
    CHECK(sh->trySetSolverValues(0, a0, 3)); CHECK(sh->trySendAssignment(0) == 0);
    int a1[] = {lit(0), lit(1, true), lit(2, true)};
    CHECK(sh->trySetSolverValues(0, a1, 3)); CHECK(sh->trySendAssignment(0) == 1);
    int b0[] = {lit(0), lit(1, true), lit(2)};
    CHECK(sh->trySetSolverValues(1, b0, 3)); CHECK(sh->trySendAssignment(1) == 0);
    sh->gpuRun(); sh->gpuRun();
    CHECK(popCount(*sh, 0) == 3); CHECK(popCount(*sh, 1) == 1); CHECK(popCount(*sh, 2) == 0);
    int c0[] = {lit(1)};
    CHECK(sh->trySetSolverValues(0, c0, 1)); sh->trySendAssignment(0);
    sh->gpuRun(); sh->gpuRun();
    CHECK(popCount(*sh, 0) == 1); CHECK(popCount(*sh, 1) == 0);
    CHECK(sh->getGlobalStat(gpuClauses) == 3);
    CHECK(std::string(sh->getGlobalStatName(gpuReports)) == "gpuReports");
    CHECK(sh->getOneSolverStat(0, reportedClauses) == 4);
    uint8_t cur[3];
    sh->getCurrentAssignment(0, cur);
    CHECK(cur[0] == 0 && cur[1] == 0 && cur[2] == 1);
    printf("shim selftest ok\n");
    return 0;
}
