// stats.h -- statistic indices and names.  Order and spelling are the reference's X-macro lists
// (gpuShareLib/GlobalStats.h:23-38, OneSolverStats.h:24-29) because consumers index by the
// enum value and print by name.
#pragma once

namespace gss {

enum GlobalStat {
    G_gpuClauses = 0,
    G_gpuClauseLengthSum,
    G_gpuClausesAdded,
    G_gpuRuns,
    G_clauseTestsOnGroups,
    G_clauseTestsOnAssigs,
    G_totalAssigClauseTested,
    G_gpuReduceDbs,
    G_gpuReports,
    G_timeSpentTestingClauses,
    G_timeSpentFillingAssigs,
    G_timeSpentFillingReported,
    G_timeSpentReduceGpuDb,
    G_COUNT
};

static const char *const kGlobalStatNames[G_COUNT] = {
    "gpuClauses", "gpuClauseLengthSum", "gpuClausesAdded", "gpuRuns", "clauseTestsOnGroups",
    "clauseTestsOnAssigs", "totalAssigClauseTested", "gpuReduceDbs", "gpuReports",
    "timeSpentTestingClauses", "timeSpentFillingAssigs", "timeSpentFillingReported",
    "timeSpentReduceGpuDb"};

enum OneSolverStat {
    S_varUpdatesSentToGpu = 0,
    S_assigsSentToGpu,
    S_failuresToFindAssig,
    S_reportedClauses,
    S_reportedClausesUnit,
    S_reportedClausesBinary,
    S_COUNT
};

static const char *const kOneSolverStatNames[S_COUNT] = {
    "varUpdatesSentToGpu", "assigsSentToGpu", "failuresToFindAssig",
    "reportedClauses", "reportedClausesUnit", "reportedClausesBinary"};

} // namespace gss
