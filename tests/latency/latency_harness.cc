// latency_harness.cc -- end-to-end import latency through the C ABI, BASELINE config 5's shape:
// N solver threads (default 64) exporting assignments and importing reported clauses, one GPU
// thread looping gss_gpu_run(), a clause database with long clauses (5 % of 101..200 literals,
// GPUSHARE_MAX_CLAUSE_LEN lifted to 200).  A probe = a fresh clause that is false under one solver's
// running assignment; its latency = gss_add_clause() ... gss_pop_reported_clause() returning it on
// that solver's thread (reference path: Solver.cc:1498-1499 sendClauseToGpu -> two gpuRun() calls ->
// GpuHelpedSolver.cc:71-92 gpuImportClauses; rel-newtech/core/Solver.cc:2891-2910).  No Python, no GIL.
// TEST / BENCH INFRASTRUCTURE: links only include/gpushare_b200.h.
//
// usage: latency_harness [solvers] [vars] [clauses] [probes] [minGpuLatencyMicros]   -> one JSON line
#include "../../include/gpushare_b200.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

using Clock = std::chrono::steady_clock;
static inline double usSince(Clock::time_point t0) { return std::chrono::duration<double, std::micro>(Clock::now() - t0).count(); }

struct Probe {
    std::atomic<int64_t> id{-1};      // clause id the target solver waits for
    std::atomic<int> target{-1};
    std::atomic<int> done{0};
    Clock::time_point t0;
    double latencyUs = 0;
};

int main(int argc, char **argv) {
    const int S = argc > 1 ? atoi(argv[1]) : 64;
    const int V = argc > 2 ? atoi(argv[2]) : 200000;
    const int64_t C = argc > 3 ? atoll(argv[3]) : 1000000;
    const int nProbes = argc > 4 ? atoi(argv[4]) : 300;
    const int minLat = argc > 5 ? atoi(argv[5]) : -1;
    // share (per mille) of the database's literals that are TRUE under the solvers' assignments: 999 = the solvers
    // satisfy nearly every learned clause (a handful of background hits per run: the latency of an otherwise
    // quiet pipeline); 985 = ~65 k background hits per run with 64 solvers (the pipeline saturated by reports)
    const int agreePerMille = argc > 6 ? atoi(argv[6]) : 999;
    const int kStable = 4096; // variables [0, kStable) are never unset by the solver threads: probes use them
    if (nProbes + 100 > kStable / 2) {
        fprintf(stderr, "at most %d probes\n", kStable / 2);
        return 2;
    }

    gss_options o;
    gss_options_default(&o);
    o.verbosity = 0;
    o.minGpuLatencyMicros = minLat;
    gss_sharer *h = gss_create(&o, nullptr, nullptr);
    gss_set_max_clause_len(h, 200);
    gss_set_var_count(h, V);
    gss_set_cpu_solver_count(h, S);

    // planted assignment + clauses that mostly agree with it (few background hits), 5 % long clauses
    std::mt19937_64 rng(12345);
    std::vector<uint8_t> sigma(V);
    for (auto &x : sigma) x = rng() & 1;
    {
        std::vector<int64_t> off(1, 0);
        std::vector<int> lits;
        lits.reserve((size_t)C * 12);
        for (int64_t c = 0; c < C; c++) {
            int len;
            const unsigned r = rng() % 100;
            if (r < 5) len = 101 + (int)(rng() % 100);
            else if (r < 55) len = 2;
            else if (r < 80) len = 3;
            else len = 4 + (int)(rng() % 27);
            for (int i = 0; i < len; i++) {
                const int v = (int)(rng() % V);
                const bool agree = (int)(rng() % 1000) < agreePerMille;
                const int sign = agree ? sigma[v] : 1 - sigma[v]; // literal true under sigma iff sign == sigma (sign 1 = negated)
                lits.push_back(2 * v + sign);
            }
            off.push_back((int64_t)lits.size());
        }
        gss_add_clauses_bulk(h, off.data(), lits.data(), C);
    }

    std::atomic<bool> stop{false};
    std::atomic<int64_t> runs{0};
    // A solver that imports a clause stops falsifying it (it backtracks).  Probe p therefore contains the false
    // literal of variable p, and once the probe is over every solver thread flips variable p: the clause is
    // satisfied from then on and does not come back run after run as a background hit.
    std::atomic<int> flipUpTo{0};
    std::vector<std::atomic<int>> flipsDone(S); // per solver thread: flips applied (and exported) so far
    for (auto &f : flipsDone) f.store(0);
    Probe probe;
    std::vector<std::thread> threads;
    std::atomic<int> ready{0};
    for (int s = 0; s < S; s++)
        threads.emplace_back([&, s] {
            std::mt19937_64 r(1000 + s);
            std::vector<int> set;
            set.reserve(V);
            for (int v = 0; v < V; v++) set.push_back(2 * v + sigma[v]); // var = sigma value (sign 1 -> false)
            while (!gss_try_set_solver_values(h, s, set.data(), (int)set.size())) std::this_thread::yield();
            while (gss_try_send_assignment(h, s) < 0) std::this_thread::yield();
            ready.fetch_add(1);
            std::vector<int> flip, fix, churn;
            int myFlips = 0;
            while (!stop.load(std::memory_order_relaxed)) {
                const int upTo = flipUpTo.load(std::memory_order_acquire);
                if (myFlips < upTo) {
                    flip.clear();
                    fix.clear();
                    for (int v = myFlips; v < upTo; v++) {
                        flip.push_back(2 * v + sigma[v]);
                        fix.push_back(2 * v + (1 - sigma[v]));
                    }
                    gss_unset_solver_values(h, s, flip.data(), (int)flip.size());
                    if (gss_try_set_solver_values(h, s, fix.data(), (int)fix.size())) myFlips = upTo;
                }
                const int flipsBefore = myFlips;
                // a little trail churn: unset and re-set a handful of (non-stable) variables, then export.  When no
                // assignment slot is free the unset is buffered and the set fails: the SAME variables are set again in the
                // next round (picking new ones instead would leave eight more variables undefined for good each time --
                // tens of thousands after a second, and every binary clause with one of them and a false literal a unit
                // hit of this solver, run after run)
                if (churn.empty()) {
                    for (int k = 0; k < 8; k++) {
                        const int v = kStable + (int)(r() % (V - kStable));
                        churn.push_back(2 * v + sigma[v]);
                    }
                    gss_unset_solver_values(h, s, churn.data(), (int)churn.size());
                }
                if (gss_try_set_solver_values(h, s, churn.data(), (int)churn.size())) {
                    churn.clear();
                    if (gss_try_send_assignment(h, s) >= 0)
                        flipsDone[s].store(flipsBefore, std::memory_order_release); // an assignment with those flips is on its way
                }
                int *lits, count;
                int64_t id;
                while (gss_pop_reported_clause(h, s, &lits, &count, &id)) {
                    if (probe.target.load(std::memory_order_acquire) == s && id == probe.id.load(std::memory_order_acquire)) {
                        probe.latencyUs = usSince(probe.t0);
                        probe.done.store(1, std::memory_order_release);
                    }
                }
                std::this_thread::sleep_for(std::chrono::microseconds(20));
            }
        });
    // device phase times and bytes of the runs, accumulated by the GPU thread (the hooks are GPU-thread only)
    double devSum[4] = {0, 0, 0, 0};
    int64_t h2dSum = 0, d2hSum = 0, sampled = 0;
    std::atomic<bool> sampling{false};
    std::thread gpu([&] {
        while (!stop.load(std::memory_order_relaxed)) {
            gss_gpu_run(h);
            runs.fetch_add(1);
            if (sampling.load(std::memory_order_relaxed)) {
                double t[4];
                int64_t a = 0, b = 0;
                if (gss_debug_last_run_times(h, t)) {
                    for (int i = 0; i < 4; i++) devSum[i] += t[i];
                    gss_debug_last_run_bytes(h, &a, &b);
                    h2dSum += a;
                    d2hSum += b;
                    sampled++;
                }
            }
        }
    });
    while (ready.load() < S) std::this_thread::sleep_for(std::chrono::milliseconds(1));
    std::this_thread::sleep_for(std::chrono::milliseconds(300)); // first runs: table rebuild, buffer growth

    std::vector<double> lat;
    int lost = 0, lostLong = 0, lostByLen[16] = {0};

    // The first kWarm probes are not measured: they take the library through its start-up transients (buffers that grow
    // to the workload's hit counts, result buffers of the right size in the pool, arenas of every probe length).
    const int kWarm = 100;
    auto tStart = Clock::now();
    int64_t runs0 = runs.load();
    double ph0[6];
    gss_debug_host_phases(h, ph0);
    int64_t reports0 = gss_get_global_stat(h, 8);
    for (int p = 0; p < nProbes + kWarm; p++) {
        if (p == kWarm) {
            tStart = Clock::now();
            runs0 = runs.load();
            gss_debug_host_phases(h, ph0);
            reports0 = gss_get_global_stat(h, 8);
            sampling.store(true);
        }
        const int s = (int)(rng() % S);
        int len = (p % 4 == 3) ? 101 + (int)(rng() % 100) : 2 + (int)(rng() % 8);
        std::vector<int> lits;
        lits.push_back(2 * p + (1 - sigma[p])); // the variable the solvers flip once the probe is over
        for (int i = 1; i < len; i++) {
            const int v = nProbes + kWarm + (int)(rng() % (kStable - nProbes - kWarm));
            lits.push_back(2 * v + (1 - sigma[v])); // false under every solver's assignment
        }
        probe.done.store(0);
        probe.id.store(-1);
        probe.target.store(s, std::memory_order_release);
        probe.t0 = Clock::now();
        const int64_t id = gss_add_clause(h, -1, lits.data(), len);
        probe.id.store(id, std::memory_order_release);
        const auto w0 = Clock::now();
        while (!probe.done.load(std::memory_order_acquire) && usSince(w0) < 1e6) std::this_thread::yield();
        if (p < kWarm) {
            // (not measured)
        } else if (probe.done.load()) lat.push_back(probe.latencyUs);
        else {
            lost++;
            if (len > 100) lostLong++;
            if (len < 16) lostByLen[len]++;
        }
        probe.target.store(-1);
        flipUpTo.store(p + 1, std::memory_order_release);
        // the next probe starts once every solver has exported an assignment that satisfies this one (bounded wait):
        // otherwise the probes of the last milliseconds come back as background hits of every solver, run after run
        const auto f0 = Clock::now();
        for (int t = 0; t < S; t++)
            while (flipsDone[t].load(std::memory_order_acquire) < p + 1 && usSince(f0) < 50e3) std::this_thread::yield();
        std::this_thread::sleep_for(std::chrono::microseconds(200));
    }
    const double wallS = usSince(tStart) * 1e-6;
    const int64_t nRuns = runs.load() - runs0;
    double ph1[6];
    gss_debug_host_phases(h, ph1);
    stop.store(true);
    for (auto &t : threads) t.join();
    gpu.join();
    const double ns = sampled ? (double)sampled : 1.0;
    std::sort(lat.begin(), lat.end());
    auto q = [&](double f) { return lat.empty() ? -1.0 : lat[std::min(lat.size() - 1, (size_t)(f * lat.size()))]; };
    printf("{\"harness\": \"import_latency\", \"solvers\": %d, \"vars\": %d, \"clauses\": %lld, \"probes\": %d, \"lost\": %d, \"lost_long\": %d, \"lost_by_len_2_to_9\": [%d,%d,%d,%d,%d,%d,%d,%d], \"true_literals_per_mille\": %d, "
           "\"p50_us\": %.1f, \"p90_us\": %.1f, \"p99_us\": %.1f, \"max_us\": %.1f, \"gpu_runs_per_s\": %.0f, "
           "\"min_gpu_latency_micros\": %d, \"long_clause_share\": 0.05, \"max_clause_len\": 200, "
           "\"host_us_per_run\": {\"finish_previous\": %.1f, \"start_next\": %.1f, \"hand_over\": %.1f, \"collect\": %.1f, "
           "\"wait_for_gpu\": %.1f}, \"hits_reported_per_run\": %.0f, "
           "\"device_us_per_run\": {\"copies\": %.1f, \"table_kernels\": %.1f, \"check_and_emit\": %.1f, \"total\": %.1f}, "
           "\"h2d_bytes_per_run\": %.0f, \"d2h_bytes_per_run\": %.0f}\n",
           S, V, (long long)C, nProbes, lost, lostLong, lostByLen[2], lostByLen[3], lostByLen[4], lostByLen[5], lostByLen[6], lostByLen[7], lostByLen[8],
           lostByLen[9], agreePerMille, q(0.5), q(0.9), q(0.99), lat.empty() ? -1.0 : lat.back(), nRuns / wallS, minLat,
           (ph1[0] - ph0[0]) / nRuns, (ph1[1] - ph0[1]) / nRuns, (ph1[2] - ph0[2]) / nRuns, (ph1[3] - ph0[3]) / nRuns,
           (ph1[4] - ph0[4]) / nRuns, (double)(gss_get_global_stat(h, 8) - reports0) / nRuns, devSum[0] / ns, devSum[1] / ns,
           devSum[2] / ns, devSum[3] / ns, h2dSum / ns, d2hSum / ns);
    gss_destroy(h);
    return 0;
}
