// common.h -- shared types and helpers of the B200 clause sharer (host + device).
#pragma once
#include <cuda_runtime.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <sstream>
#include <string>
#include <sys/resource.h>

namespace gss {

// ---- literal encoding: same as the reference (gpuShareLib/SolverTypes.h:44-60) ----
// lit = 2*var + sign, sign 1 = negated
__host__ __device__ inline int litVar(int lit) { return lit >> 1; }
__host__ __device__ inline int litSign(int lit) { return lit & 1; }

// value of a variable, same numbering as the reference's lbool (SolverTypes.h:77-80)
enum : uint8_t { V_TRUE = 0, V_FALSE = 1, V_UNDEF = 2 };

constexpr int kSlots = 32;          // assignment slots per solver (one 32-bit word)
constexpr int kTileClauses = 128;   // clauses per tile: one warp x int4
// Position of clause j (0..127) of a tile inside each of the tile's literal rows.  Lane l of the warp
// that checks the tile loads words 4l..4l+3 of a row with one 128-bit load and must find there the
// clauses l, l+32, l+64, l+96: then each of its four level-1 gathers is issued for 32 CONSECUTIVE
// clauses, whose first literals are neighbours in the (first-literal-sorted) arena -- a handful of
// 32-byte sectors per gather instead of the whole tile's range four times over.
__host__ __device__ inline int tileSlot(int j) { return ((j & 31) << 2) | (j >> 5); }
constexpr int kDefaultMaxClauseLen = 100; // reference MAX_CL_SIZE (BaseTypes.cuh:28)
constexpr int kMaxSupportedClauseLen = 65535; // Survivor::ptrLen keeps the length in 16 bits
constexpr int kMaxSolversPerGroup = 32;   // one aggregate word / one lane per solver

// One delta record: the 32 slots of one variable of one solver (reference VarUpdate,
// gpuShareLib/Assigs.cuh:122-125).  tru is garbage where def is 0.
struct VarUpdate {
    int32_t var;
    uint32_t def;
    uint32_t tru;
};

// What the check kernels emit (reference ReportedClause, BaseTypes.cuh:146-151)
struct HitRecord {
    uint32_t mask;
    int32_t solver;
    int32_t len;
    int32_t idx; // index of the clause inside its length array
};

// A hit after the GPU-side post-processing of a large result list: ordered by (solver, length,
// index), with the clause id and the position of the clause's literals in the emitted literal stream
struct SortedHit {
    uint32_t mask;
    int32_t solver;
    int32_t len;
    int32_t idx;
    int64_t id;
    int64_t litPos;
};

// A clause that survived the aggregate filter
struct Survivor {
    uint64_t ptrLen; // device address of the clause's first literal (low 48 bits) | length << 48
    int32_t idx;     // index of the clause inside its length array
    uint32_t aggBits;
};

// Per-solver parameters of one run (reference DOneSolverAssigs + AggCorresp,
// gpuShareLib/Assigs.cuh:77-80,105-111)
struct SolverRunParams {
    uint32_t startVals;  // frozen slots = the ones to test
    uint32_t lastMask;   // the slot every other slot collapses to after the run
    uint32_t allAggBits; // aggregate bits owned by this solver
    uint32_t usedAggBits; // aggregate bits in use this run
    int32_t updStart;    // range of this solver's updates in the update array
    int32_t updCount;
    int32_t nGroups;     // number of (aggregate bit, slot mask) pairs
    int32_t pad;
    uint32_t groupAggBit[kSlots];
    uint32_t groupSlotMask[kSlots];
};

struct Logger {
    int verbosity = 1;
    std::function<void(const std::string &)> fn;
    void log(int level, const std::string &s) const {
        if (level <= verbosity && fn) fn(s);
    }
};

[[noreturn]] inline void die(const char *file, int line, const std::string &msg) {
    // same contract as the reference's THROW() (gpuShareLib/Assert.h:44-48): print and exit(1)
    fprintf(stderr, "gpushare_b200 fatal: %s (%s:%d)\n", msg.c_str(), file, line);
    fflush(stderr);
    exit(1);
}

#define GSS_DIE(msg) ::gss::die(__FILE__, __LINE__, (msg))
#define GSS_CHECK(cond)                                        \
    do {                                                       \
        if (!(cond)) GSS_DIE(std::string("check failed: ") + #cond); \
    } while (0)
#define GSS_CUDA(call)                                                            \
    do {                                                                          \
        cudaError_t e__ = (call);                                                 \
        if (e__ != cudaSuccess)                                                   \
            GSS_DIE(std::string("CUDA error ") + cudaGetErrorString(e__) + " in " + #call); \
    } while (0)

// Fine-grained host profile of the run pipeline (GSS_HOST_PROF=1: mean microseconds per call of every
// named scope, printed to stderr when the process exits).  Off: one predictable branch per scope.
struct HostProf {
    static constexpr int kMax = 32;
    struct Entry {
        const char *name;
        double us, maxUs;
        long calls, ctxSwitches, pageFaults; // (of the calling thread, inside the scope: getrusage)
    };
    static bool enabled() {
        static const bool on = getenv("GSS_HOST_PROF") != nullptr;
        return on;
    }
    static Entry *table() {
        static Entry t[kMax] = {};
        static const bool registered = (atexit(&HostProf::dump), true);
        (void)registered;
        return t;
    }
    static void add(const char *name, double us, long ctx, long faults) {
        Entry *t = table();
        for (int i = 0; i < kMax; i++) {
            if (!t[i].name) t[i].name = name;
            if (t[i].name == name) {
                t[i].us += us;
                t[i].maxUs = us > t[i].maxUs ? us : t[i].maxUs;
                t[i].calls++;
                t[i].ctxSwitches += ctx;
                t[i].pageFaults += faults;
                return;
            }
        }
    }
    static void reset() { // (bench: gss_debug_host_phases marks the start of a timed region)
        Entry *t = table();
        for (int i = 0; i < kMax; i++) t[i].us = t[i].maxUs = 0, t[i].calls = t[i].ctxSwitches = t[i].pageFaults = 0;
    }
    static void dump() {
        if (!enabled()) return;
        Entry *t = table();
        for (int i = 0; i < kMax && t[i].name; i++)
            if (t[i].calls)
                fprintf(stderr, "host prof %-28s %9.2f us/call x %ld  (max %.1f us; %ld context switches, %ld page faults)\n", t[i].name,
                        t[i].us / (double)t[i].calls, t[i].calls, t[i].maxUs, t[i].ctxSwitches, t[i].pageFaults);
    }
    const char *name;
    std::chrono::steady_clock::time_point t0;
    long ctx0 = 0, flt0 = 0;
    static void usage(long &ctx, long &flt) {
        struct rusage ru;
        getrusage(RUSAGE_THREAD, &ru);
        ctx = ru.ru_nivcsw + ru.ru_nvcsw;
        flt = ru.ru_minflt + ru.ru_majflt;
    }
    explicit HostProf(const char *n) : name(n) {
        if (enabled()) {
            usage(ctx0, flt0);
            t0 = std::chrono::steady_clock::now();
        }
    }
    ~HostProf() {
        if (!enabled()) return;
        const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        long ctx, flt;
        usage(ctx, flt);
        add(name, us, ctx - ctx0, flt - flt0);
    }
};

} // namespace gss
