// synth.cc -- deterministic synthetic inputs (see gss_synth.h).  TEST / BENCH INFRASTRUCTURE: its own
// host-only library (libgss_synth.so), so that input generation never loads the product library.
#include "gss_synth.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

namespace {

struct SplitMix64 {
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    // uniform in [0, n)
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
    double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

// Glucose::luby(2, x) restated (glucose-syrup/core/Solver.cc:1967-1979): 2^seq
inline int lubyPow2(int64_t x) {
    int64_t size = 1;
    int seq = 0;
    while (size < x + 1) { seq++; size = 2 * size + 1; }
    while (size - 1 != x) {
        size = (size - 1) >> 1;
        seq--;
        x = x % size;
    }
    return 1 << seq;
}

inline int clauseLen(int64_t i, int maxLen) {
    int l = 1 + lubyPow2(i);
    return l < maxLen ? l : maxLen;
}

} // namespace

struct gss_synth_stream {
    int nvars;
    std::vector<uint8_t> sigma, vals;
    double pUndef;
    int perStep;
    SplitMix64 rng;
    bool first = true;
    std::vector<int> touched;
    std::vector<uint8_t> mark;
    gss_synth_stream(int n, uint64_t seed) : nvars(n), rng(seed) {}
};

extern "C" {

int64_t gss_synth_total_lits(int64_t nclauses, int max_len) {
    int64_t t = 0;
    for (int64_t i = 0; i < nclauses; i++) t += clauseLen(i, max_len);
    return t;
}

void gss_synth_sigma(int nvars, uint64_t seed, uint8_t *sigma) {
    SplitMix64 r(seed);
    for (int v = 0; v < nvars; v++) sigma[v] = (uint8_t)(r.next() & 1);
}

void gss_synth_clauses(int64_t nclauses, int nvars, int max_len, const uint8_t *sigma, double p_agree,
                       uint64_t seed, int64_t *offsets, int32_t *lits) {
    // Every literal consumes exactly two draws of ONE SplitMix64 stream (variable, then agreement / sign)
    // and SplitMix64 jumps ahead in O(1) (its state is seed + k * gamma): clause ranges are generated
    // concurrently and the output is bit-identical to the sequential stream, whatever the thread count
    // (GSS_SYNTH_THREADS overrides it; tests compare 1 thread against many).
    int64_t pos = 0;
    for (int64_t c = 0; c < nclauses; c++) {
        offsets[c] = pos;
        pos += clauseLen(c, max_len);
    }
    offsets[nclauses] = pos;
    int nThreads = (int)std::thread::hardware_concurrency();
    if (const char *e = getenv("GSS_SYNTH_THREADS")) nThreads = atoi(e);
    if (nThreads < 1) nThreads = 1;
    if (nclauses < 100000) nThreads = 1;
    auto range = [&](int64_t c0, int64_t c1) {
        SplitMix64 r(seed + 2ull * (uint64_t)offsets[c0] * 0x9E3779B97F4A7C15ull);
        for (int64_t c = c0; c < c1; c++) {
            int32_t *out = lits + offsets[c];
            const int len = (int)(offsets[c + 1] - offsets[c]);
            for (int i = 0; i < len; i++) {
                int v = (int)r.below((uint32_t)nvars);
                int sign;
                if (sigma) {
                    bool agree = r.unit() < p_agree;
                    // literal true under sigma: positive when sigma says true (0), negated otherwise
                    int trueSign = sigma[v] == 0 ? 0 : 1;
                    sign = agree ? trueSign : 1 - trueSign;
                } else {
                    sign = (int)(r.next() & 1);
                }
                out[i] = 2 * v + sign;
            }
        }
    };
    if (nThreads == 1) {
        range(0, nclauses);
        return;
    }
    std::vector<std::thread> threads;
    for (int t = 0; t < nThreads; t++) {
        const int64_t c0 = nclauses * t / nThreads, c1 = nclauses * (t + 1) / nThreads;
        if (c1 > c0) threads.emplace_back(range, c0, c1);
    }
    for (auto &th : threads) th.join();
}

gss_synth_stream *gss_synth_stream_create(int nvars, const uint8_t *sigma, double p_undef, double churn, uint64_t seed) {
    gss_synth_stream *s = new gss_synth_stream(nvars, seed);
    s->sigma.assign(sigma, sigma + nvars);
    s->vals.assign(nvars, 2);
    s->pUndef = p_undef;
    s->perStep = (int)std::ceil(churn * nvars);
    s->mark.assign(nvars, 0);
    return s;
}

void gss_synth_stream_destroy(gss_synth_stream *s) { delete s; }

const uint8_t *gss_synth_stream_values(gss_synth_stream *s) { return s->vals.data(); }

void gss_synth_stream_next(gss_synth_stream *s, int32_t *set_lits, int32_t *n_set, int32_t *unset_lits, int32_t *n_unset) {
    int ns = 0, nu = 0;
    if (s->first) {
        s->first = false;
        for (int v = 0; v < s->nvars; v++) {
            uint8_t nv = s->rng.unit() < s->pUndef ? 2 : s->sigma[v];
            s->vals[v] = nv;
            if (nv != 2) set_lits[ns++] = 2 * v + (nv == 1 ? 1 : 0);
        }
    } else {
        s->touched.clear();
        for (int k = 0; k < s->perStep; k++) {
            int v = (int)s->rng.below((uint32_t)s->nvars);
            uint8_t nv = s->rng.unit() < s->pUndef ? 2 : s->sigma[v];
            if (s->mark[v]) continue; // one change per variable per step
            if (nv == s->vals[v]) continue;
            s->mark[v] = 1;
            s->touched.push_back(v);
            s->vals[v] = nv;
            if (nv == 2) unset_lits[nu++] = 2 * v; // the sign is irrelevant for an unset
            else set_lits[ns++] = 2 * v + (nv == 1 ? 1 : 0);
        }
        for (int v : s->touched) s->mark[v] = 0;
    }
    *n_set = ns;
    *n_unset = nu;
}


int gss_synth_write_cnf(const char *path, int nvars, int64_t nclauses, const double *len_weights, int n_lens,
                        double p_agree, uint64_t seed) {
    FILE *f = fopen(path, "w");
    if (!f) return -1;
    std::vector<uint8_t> sigma((size_t)nvars);
    gss_synth_sigma(nvars, seed ^ 0x5151ull, sigma.data());
    double total = 0;
    for (int k = 0; k < n_lens; k++) total += len_weights[k];
    SplitMix64 r(seed);
    std::vector<char> buf;
    buf.reserve(1 << 22);
    fprintf(f, "p cnf %d %lld\n", nvars, (long long)nclauses);
    int vars[64];
    for (int64_t c = 0; c < nclauses; c++) {
        double x = r.unit() * total;
        int len = 2;
        for (int k = 0; k < n_lens; k++) {
            if (x < len_weights[k] || k == n_lens - 1) { len = k + 2; break; }
            x -= len_weights[k];
        }
        if (len > 64) len = 64;
        bool sat = false;
        int signs[64];
        for (int i = 0; i < len; i++) {
            int v;
            bool dup;
            do {
                v = (int)r.below((uint32_t)nvars);
                dup = false;
                for (int j = 0; j < i; j++) dup = dup || vars[j] == v;
            } while (dup);
            vars[i] = v;
            const bool agree = r.unit() < p_agree;
            const int trueSign = sigma[(size_t)v] == 0 ? 0 : 1;
            signs[i] = agree ? trueSign : 1 - trueSign;
            sat = sat || agree;
        }
        if (!sat) {
            const int i = (int)r.below((uint32_t)len);
            signs[i] = 1 - signs[i];
        }
        char tmp[16];
        for (int i = 0; i < len; i++) {
            int n = snprintf(tmp, sizeof(tmp), "%s%d ", signs[i] ? "-" : "", vars[i] + 1);
            buf.insert(buf.end(), tmp, tmp + n);
        }
        buf.push_back('0');
        buf.push_back('\n');
        if (buf.size() > (1u << 22) - 1024) {
            fwrite(buf.data(), 1, buf.size(), f);
            buf.clear();
        }
    }
    fwrite(buf.data(), 1, buf.size(), f);
    return fclose(f) == 0 ? 0 : -1;
}
} // extern "C"
