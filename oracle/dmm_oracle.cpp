// dmm_oracle.cpp — CPU ORACLE. TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE.
//
// A scalar, single-threaded-per-replica restatement of the digital-memcomputing ODE
// integrator of AHartNtkn/odesat (`src/system.rs`), operation for operation, templated on
// float/double.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` leg may load this library; the product (odesat_b200/) never does.
//
// PARITY UNPINNED BY THE REFERENCE: the reference has no tests, golden vectors or KATs
//  and no Rust toolchain exists in this image, so the reference binary
// cannot be run.  The oracle is pinned instead against the hand-derived KATs of SURVEY.md
// §8c on tests/small.cnf (tests/test_oracle_kat.py) and against an independent pure-Python
// restatement (oracle/pyref.py).
//
// Build: g++ -O2 -ffp-contract=off (no FMA contraction: Rust never contracts).
//
// Every function cites the reference lines it follows (paths relative to /root/reference).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

struct Formula {
    int64_t varnum = 0;              // cnf.rs:56 (header value, may exceed distinct vars)
    int64_t n_clauses = 0;
    std::vector<int64_t> off;        // [M+1]
    std::vector<int32_t> var;        // [L] 0-based variable index  (cnf.rs:7)
    std::vector<uint8_t> neg;        // [L] is_negated              (cnf.rs:8)
};

// system.rs:19-23
template <typename T> struct K {
    static constexpr T ALPHA = T(5.0);
    static constexpr T BETA = T(20.0);
    static constexpr T GAMMA = T(0.25);
    static constexpr T DELTA = T(0.05);
    static constexpr T EPSILON = T(0.001);
};

// Rust f64::max / f64::min ignore a NaN operand — the same contract as C fmax/fmin.
// Written out (no libm call): NaN in one operand returns the other one.
template <typename T> inline T rmax(T a, T b) { return (b != b) ? a : ((a != a) ? b : (a > b ? a : b)); }
template <typename T> inline T rmin(T a, T b) { return (b != b) ? a : ((a != a) ? b : (a < b ? a : b)); }

// system.rs:25-91 compute_derivatives
template <typename T>
bool compute_derivatives(const Formula& f, const T* v, const T* xs, const T* xl, T zeta, T* dv,
                         T* dxs, T* dxl) {
    const T INF = std::numeric_limits<T>::infinity();
    for (int64_t i = 0; i < f.varnum; ++i) dv[i] = T(0);                  // :33
    bool all = true;                                                       // :90 fold(true, &&)
    T val[64];
    std::vector<T> big;
    for (int64_t m = 0; m < f.n_clauses; ++m) {                            // :35-40
        const int64_t b = f.off[m], e = f.off[m + 1];
        T* vals = val;
        if (e - b > 64) { big.resize(e - b); vals = big.data(); }
        T mn = INF, sm = INF;                                              // :43-44
        for (int64_t j = b; j < e; ++j) {                                  // :46-57
            const T q = f.neg[j] ? T(-1) : T(1);                           // :47
            const T vi = v[f.var[j]];                                      // :48
            const T value = T(1) - q * vi;                                 // :49
            if (value < mn) { sm = mn; mn = value; }                       // :50-52
            else if (value < sm) { sm = value; }                           // :53-55
            vals[j - b] = value;                                           // :56
        }
        const T c_m = T(0.5) * mn;                                         // :60
        const T xs_m = xs[m], xl_m = xl[m];
        for (int64_t j = b; j < e; ++j) {                                  // :62
            const T q = f.neg[j] ? T(-1) : T(1);
            const int32_t i = f.var[j];
            const T value = vals[j - b];
            const T g = T(0.5) * q * ((value != mn) ? mn : sm);            // :64-70
            const T r = (c_m == (T(1) - q * v[i])) ? T(0.5) * (q - v[i]) : T(0);   // :73-77
            dv[i] += xl_m * xs_m * g + (T(1) + zeta * xl_m) * (T(1) - xs_m) * r;   // :80
        }
        dxs[m] = K<T>::BETA * (xs_m + K<T>::EPSILON) * (c_m - K<T>::GAMMA);        // :84
        dxl[m] = K<T>::ALPHA * (c_m - K<T>::DELTA);                                 // :85
        all = all && (c_m < K<T>::GAMMA);                                           // :88, :90
    }
    return all;
}

// system.rs:93-97 update_state
template <typename T>
void update_state(const Formula& f, T* v, T* xs, T* xl, const T* dv, const T* dxs, const T* dxl,
                  T dt) {
    const T hi_s = T(1) - K<T>::EPSILON;
    const T hi_l = T(1e4) * T(f.n_clauses);
    for (int64_t m = 0; m < f.n_clauses; ++m)
        xs[m] = rmin(rmax(xs[m] + dt * dxs[m], K<T>::EPSILON), hi_s);     // :94
    for (int64_t m = 0; m < f.n_clauses; ++m)
        xl[m] = rmin(rmax(xl[m] + dt * dxl[m], T(1)), hi_l);              // :95
    for (int64_t i = 0; i < f.varnum; ++i)
        v[i] = rmin(rmax(v[i] + dt * dv[i], T(-1)), T(1));                // :96
}

// system.rs:101-109 max_error (folds start at NaN; max ignores NaN)
template <typename T>
T max_error(int64_t N, int64_t M, const T* av, const T* axs, const T* axl, const T* bv,
            const T* bxs, const T* bxl) {
    const T NANV = std::numeric_limits<T>::quiet_NaN();
    T ev = NANV, es = NANV, el = NANV;
    for (int64_t i = 0; i < N; ++i) ev = rmax(ev, std::fabs(av[i] - bv[i]));
    for (int64_t m = 0; m < M; ++m) es = rmax(es, std::fabs(axs[m] - bxs[m]));
    for (int64_t m = 0; m < M; ++m) el = rmax(el, std::fabs(axl[m] - bxl[m]));
    return rmax(ev, rmax(es, el));
}

template <typename T> struct Scratch {
    std::vector<T> dv, dxs, dxl, tv, txs, txl;
    explicit Scratch(const Formula& f)
        : dv(f.varnum), dxs(f.n_clauses), dxl(f.n_clauses), tv(f.varnum), txs(f.n_clauses),
          txl(f.n_clauses) {}
};

// system.rs:141-154 euler_step_fixed — updates unconditionally, returns the pre-update flag
template <typename T>
bool euler_step_fixed(const Formula& f, T* v, T* xs, T* xl, T dt, T zeta, Scratch<T>& s) {
    const bool allsat = compute_derivatives(f, v, xs, xl, zeta, s.dv.data(), s.dxs.data(),
                                            s.dxl.data());
    update_state(f, v, xs, xl, s.dv.data(), s.dxs.data(), s.dxl.data(), dt);
    return allsat;
}

// system.rs:111-139 euler_step — one full step vs two half steps; never rejects
template <typename T>
bool euler_step(const Formula& f, T* v, T* xs, T* xl, T tol, T* dt, T zeta, Scratch<T>& s) {
    const int64_t N = f.varnum, M = f.n_clauses;
    const bool allsat = compute_derivatives(f, v, xs, xl, zeta, s.dv.data(), s.dxs.data(),
                                            s.dxl.data());                 // :120
    if (!allsat) {
        std::memcpy(s.tv.data(), v, sizeof(T) * N);                        // :124 clone
        std::memcpy(s.txs.data(), xs, sizeof(T) * M);
        std::memcpy(s.txl.data(), xl, sizeof(T) * M);
        update_state(f, s.tv.data(), s.txs.data(), s.txl.data(), s.dv.data(), s.dxs.data(),
                     s.dxl.data(), *dt);                                   // :125
        update_state(f, v, xs, xl, s.dv.data(), s.dxs.data(), s.dxl.data(), T(0.5) * *dt);  // :128
        compute_derivatives(f, v, xs, xl, zeta, s.dv.data(), s.dxs.data(), s.dxl.data());   // :129
        update_state(f, v, xs, xl, s.dv.data(), s.dxs.data(), s.dxl.data(), T(0.5) * *dt);  // :130
        const T err = max_error(N, M, s.tv.data(), s.txs.data(), s.txl.data(), (const T*)v,
                                (const T*)xs, (const T*)xl);               // :132
        *dt = rmax(rmin(*dt * std::sqrt(tol / err), T(1e3)), T(0.0078125));   // :133-135 (2^-7)
    }
    return allsat;
}

// system.rs:164-173 zeta density rule
template <typename T> T default_zeta(const Formula& f) {
    const double d = double(f.n_clauses) / double(f.varnum);
    return d >= 6.0 ? T(0.1) : (d >= 4.9 ? T(0.01) : T(0.001));
}

// system.rs:156-239 simulate.  Options as sentinels: tol NaN → 1e-3 (:174); step_size NaN →
// adaptive with dt0 = 0.01 (:205); steps < 0 → unbounded (:198, :221); zeta NaN → density rule.
template <typename T>
int simulate(const Formula& f, T* v, T* xs, T* xl, double tol_, double step_, int64_t steps,
             double zeta_, uint8_t* assign, int64_t* steps_taken, double* dt_out) {
    const T zeta = std::isnan(zeta_) ? default_zeta<T>(f) : T(zeta_);
    const T tol = std::isnan(tol_) ? T(1e-3) : T(tol_);
    Scratch<T> s(f);
    int64_t it = 0;
    bool flag = false;
    T dt = T(0.01);
    if (!std::isnan(step_)) {
        const T h = T(step_);
        dt = h;
        while (steps < 0 || it < steps) {
            ++it;
            if (euler_step_fixed(f, v, xs, xl, h, zeta, s)) { flag = true; break; }
        }
    } else {
        while (steps < 0 || it < steps) {
            ++it;
            if (euler_step(f, v, xs, xl, tol, &dt, zeta, s)) { flag = true; break; }
        }
    }
    if (assign) for (int64_t i = 0; i < f.varnum; ++i) assign[i] = v[i] > T(0);   // :238
    if (steps_taken) *steps_taken = it;
    if (dt_out) *dt_out = double(dt);
    return flag ? 1 : 0;
}

// system.rs:241-359 simulate_inter.  states laid out [R][N] / [R][M].  Adaptive mode shares
// ONE dt across replicas (:314, quirk Q7); state_res starts all-true (:274, quirk Q8).
template <typename T>
int64_t simulate_inter(const Formula& f, int64_t R, T* v, T* xs, T* xl, double tol_, double step_,
                       int64_t steps, double zeta_, uint8_t* assign, int64_t* steps_taken) {
    const int64_t N = f.varnum, M = f.n_clauses;
    const T zeta = std::isnan(zeta_) ? default_zeta<T>(f) : T(zeta_);
    const T tol = std::isnan(tol_) ? T(1e-3) : T(tol_);
    Scratch<T> s(f);
    std::vector<uint8_t> res(R, 1);                                        // :274
    int64_t it = 0;
    T dt = T(0.01);                                                        // :314
    const bool fixed = !std::isnan(step_);
    while (steps < 0 || it < steps) {
        ++it;
        for (int64_t r = 0; r < R; ++r)
            res[r] = fixed ? euler_step_fixed(f, v + r * N, xs + r * M, xl + r * M, T(step_), zeta, s)
                           : euler_step(f, v + r * N, xs + r * M, xl + r * M, tol, &dt, zeta, s);
        bool any = false;
        for (int64_t r = 0; r < R; ++r) any = any || res[r];               // :291
        if (any) break;
    }
    int64_t win = -1;
    for (int64_t r = 0; r < R; ++r) if (res[r]) { win = r; break; }        // :353
    const int64_t src = win >= 0 ? win : 0;                                // :357
    if (assign && R > 0) for (int64_t i = 0; i < N; ++i) assign[i] = v[src * N + i] > T(0);
    if (steps_taken) *steps_taken = it;
    return win;
}

// system.rs:362-372 init_short_term_memory
template <typename T> void init_short_term_memory(const Formula& f, T* xs) {
    for (int64_t m = 0; m < f.n_clauses; ++m) {
        bool anyneg = false;
        for (int64_t j = f.off[m]; j < f.off[m + 1]; ++j) anyneg = anyneg || f.neg[j];
        xs[m] = anyneg ? T(1) : T(-1);
    }
}

// cnf.rs:246-264 evaluate_cnf on a dense assignment over 0..varnum-1
int evaluate_cnf(const Formula& f, const uint8_t* a) {
    for (int64_t m = 0; m < f.n_clauses; ++m) {
        bool c = false;
        for (int64_t j = f.off[m]; j < f.off[m + 1]; ++j) {
            const bool val = a[f.var[j]] != 0;
            c = c || (f.neg[j] ? !val : val);
        }
        if (!c) return 0;
    }
    return 1;
}

// Seeded stand-in for main.rs:170-174 (`rng.gen::<f64>() * 2.0 - 1.0`, OS-seeded ChaCha in
// the reference — unreproducible by construction).  Counter-based SplitMix64 keyed by
// (seed, replica, variable); the CUDA init kernel implements the same function bit for bit.
inline uint64_t sm64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
inline uint64_t v0_key(uint64_t seed, uint64_t replica) { return sm64(seed ^ sm64(replica)); }
inline uint64_t v0_bits(uint64_t seed, uint64_t replica, uint64_t var) {
    return sm64(v0_key(seed, replica) ^ (var * 0xD1342543DE82EF95ull));
}
template <typename T> T v0_value(uint64_t bits);
template <> double v0_value<double>(uint64_t bits) {
    return double(bits >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;   // 53-bit, [-1,1)
}
template <> float v0_value<float>(uint64_t bits) {
    return float(bits >> 40) * (1.0f / 16777216.0f) * 2.0f - 1.0f;        // 24-bit, [-1,1)
}

// batch (main.rs:278-308) restated for many replicas: each replica runs `simulate` with
// Some(steps); optional freeze-free mode (freeze=0) keeps stepping past the flag so the work
// is a fixed steps×M×R (throughput baseline).  Replicas are distributed over host threads.
template <typename T>
void batch_fixed(const Formula& f, int64_t R, T* v, T* xs, T* xl, T dt, T zeta, int64_t steps,
                 int freeze, int64_t* solved_step, int nthreads) {
    const int64_t N = f.varnum, M = f.n_clauses;
    auto work = [&](int64_t r0, int64_t r1) {
        Scratch<T> s(f);
        for (int64_t r = r0; r < r1; ++r) {
            int64_t first = -1;
            for (int64_t it = 0; it < steps; ++it) {
                const bool flag = euler_step_fixed(f, v + r * N, xs + r * M, xl + r * M, dt, zeta, s);
                if (flag && first < 0) { first = it; if (freeze) break; }
            }
            if (solved_step) solved_step[r] = first;
        }
    };
    if (nthreads <= 1) { work(0, R); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        const int64_t r0 = R * t / nthreads, r1 = R * (t + 1) / nthreads;
        if (r1 > r0) th.emplace_back(work, r0, r1);
    }
    for (auto& t : th) t.join();
}

template <typename T>
void batch_adaptive(const Formula& f, int64_t R, T* v, T* xs, T* xl, T tol, T zeta, int64_t steps,
                    int64_t* solved_step, T* dt_out, int nthreads) {
    const int64_t N = f.varnum, M = f.n_clauses;
    auto work = [&](int64_t r0, int64_t r1) {
        Scratch<T> s(f);
        for (int64_t r = r0; r < r1; ++r) {
            int64_t first = -1;
            T dt = T(0.01);
            for (int64_t it = 0; it < steps; ++it)
                if (euler_step(f, v + r * N, xs + r * M, xl + r * M, tol, &dt, zeta, s)) { first = it; break; }
            if (solved_step) solved_step[r] = first;
            if (dt_out) dt_out[r] = dt;
        }
    };
    if (nthreads <= 1) { work(0, R); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        const int64_t r0 = R * t / nthreads, r1 = R * (t + 1) / nthreads;
        if (r1 > r0) th.emplace_back(work, r0, r1);
    }
    for (auto& t : th) t.join();
}

}  // namespace

#define F(p) (*static_cast<const Formula*>(p))

// ---------------------------------------------------------------------------------------------
// src/stoch.rs — the stochastic local search (a separate algorithm).
// The reference draws `rng.gen_range(1..=total)` from an OS-seeded ThreadRng (stoch.rs:68, :81), which no run can
// reproduce; the oracle and the GPU kernels share a counter-based stand-in: r = 1 + mulhi64(h, total) with
// h = SplitMix64 of (seed, replica, step, variable).  Everything else is the reference's integer arithmetic.
// ---------------------------------------------------------------------------------------------
constexpr uint64_t STOCH_ALPHA = 20;                                       // stoch.rs:18
inline uint64_t stoch_bits(uint64_t seed, uint64_t replica, uint64_t step, uint64_t var) {
    return sm64(v0_key(seed, replica) ^ sm64(step + 0x632BE59BD9B4E019ull) ^ (var * 0xD1342543DE82EF95ull));
}

// stoch.rs:26-78 step → all clauses satisfied (by the state before the flips)
bool stoch_step(const Formula& f, uint8_t* v, uint64_t* xl, uint64_t seed, uint64_t replica, uint64_t step,
                std::vector<uint64_t>& unsat_w, std::vector<uint64_t>& total_w) {
    bool all = true;                                                       // :32
    unsat_w.assign((size_t)f.varnum, 0);                                   // :35-38 slab of (0, 0)
    total_w.assign((size_t)f.varnum, 0);
    for (int64_t m = 0; m < f.n_clauses; ++m) {                            // :41-64
        bool sat = false;                                                  // :20-25 evaluate_clause
        for (int64_t j = f.off[m]; j < f.off[m + 1]; ++j) sat = sat || ((v[f.var[j]] != 0) != (f.neg[j] != 0));
        uint64_t x = xl[m];
        if (sat) { x = x > 0 ? x - 1 : 0; x = x < 1 ? 1 : x; }             // :48 saturating_sub(1).max(1)
        else { x = x > UINT64_MAX - STOCH_ALPHA ? UINT64_MAX : x + STOCH_ALPHA; }   // :50 saturating_add(ALPHA)
        xl[m] = x;
        for (int64_t j = f.off[m]; j < f.off[m + 1]; ++j) {                // :54-59
            total_w[f.var[j]] += x;
            if (!sat) unsat_w[f.var[j]] += x;
        }
        if (!sat) all = false;                                             // :61-63
    }
    for (int64_t i = 0; i < f.varnum; ++i) {                               // :67-74
        if (total_w[i] == 0) continue;   // the reference panics here (gen_range over an empty range): variable in no clause
        const uint64_t h = stoch_bits(seed, replica, step, (uint64_t)i);
        const uint64_t r = 1 + (uint64_t)(((unsigned __int128)h * total_w[i]) >> 64);
        if (r <= unsat_w[i]) v[i] = !v[i];
    }
    return all;                                                            // :77
}

// stoch.rs:80-110 search for R independent replicas: each runs until its own step returns true or `steps` are done
void stoch_batch(const Formula& f, int64_t R, uint8_t* v, uint64_t* xl, uint64_t seed, int64_t replica_offset,
                 int64_t steps, int64_t* solved_step) {
    std::vector<uint64_t> a, b;
    for (int64_t r = 0; r < R; ++r) {
        solved_step[r] = -1;
        for (int64_t s = 0; s < steps; ++s) {                              // :95-99
            if (stoch_step(f, v + r * f.varnum, xl + r * f.n_clauses, seed, (uint64_t)(replica_offset + r), (uint64_t)s, a, b)) {
                solved_step[r] = s;
                break;
            }
        }
    }
}

extern "C" {

// lits: signed DIMACS-style ±(index+1) over the NORMALISED variables 0..varnum-1.
void* dmm_formula_new(int64_t varnum, int64_t n_clauses, const int64_t* off, const int32_t* lits) {
    auto* f = new Formula;
    f->varnum = varnum;
    f->n_clauses = n_clauses;
    f->off.assign(off, off + n_clauses + 1);
    const int64_t L = off[n_clauses];
    f->var.resize(L);
    f->neg.resize(L);
    for (int64_t j = 0; j < L; ++j) {
        const int32_t l = lits[j];
        const int64_t a = l < 0 ? -int64_t(l) : int64_t(l);
        if (a < 1 || a > varnum) { delete f; return nullptr; }   // reference: index-OOB panic
        f->var[j] = int32_t(a - 1);
        f->neg[j] = l < 0;
    }
    return f;
}
void dmm_formula_free(void* f) { delete static_cast<Formula*>(f); }

#define ORACLE_API(T, SFX)                                                                        \
    int dmm_compute_derivatives_##SFX(const void* f, const T* v, const T* xs, const T* xl,        \
                                      double zeta, T* dv, T* dxs, T* dxl) {                       \
        return compute_derivatives<T>(F(f), v, xs, xl, T(zeta), dv, dxs, dxl) ? 1 : 0;            \
    }                                                                                             \
    void dmm_update_state_##SFX(const void* f, T* v, T* xs, T* xl, const T* dv, const T* dxs,     \
                                const T* dxl, double dt) {                                        \
        update_state<T>(F(f), v, xs, xl, dv, dxs, dxl, T(dt));                                    \
    }                                                                                             \
    double dmm_max_error_##SFX(int64_t N, int64_t M, const T* av, const T* axs, const T* axl,     \
                               const T* bv, const T* bxs, const T* bxl) {                         \
        return double(max_error<T>(N, M, av, axs, axl, bv, bxs, bxl));                            \
    }                                                                                             \
    int dmm_euler_step_fixed_##SFX(const void* f, T* v, T* xs, T* xl, double dt, double zeta) {   \
        Scratch<T> s(F(f));                                                                       \
        return euler_step_fixed<T>(F(f), v, xs, xl, T(dt), T(zeta), s) ? 1 : 0;                   \
    }                                                                                             \
    int dmm_euler_step_##SFX(const void* f, T* v, T* xs, T* xl, double tol, double* dt,           \
                             double zeta) {                                                       \
        Scratch<T> s(F(f));                                                                       \
        T d = T(*dt);                                                                             \
        const bool a = euler_step<T>(F(f), v, xs, xl, T(tol), &d, T(zeta), s);                    \
        *dt = double(d);                                                                          \
        return a ? 1 : 0;                                                                         \
    }                                                                                             \
    void dmm_init_short_term_memory_##SFX(const void* f, T* xs) {                                 \
        init_short_term_memory<T>(F(f), xs);                                                      \
    }                                                                                             \
    int dmm_simulate_##SFX(const void* f, T* v, T* xs, T* xl, double tol, double step_size,       \
                           int64_t steps, double zeta, uint8_t* assign, int64_t* steps_taken,     \
                           double* dt_out) {                                                      \
        return simulate<T>(F(f), v, xs, xl, tol, step_size, steps, zeta, assign, steps_taken,     \
                           dt_out);                                                               \
    }                                                                                             \
    int64_t dmm_simulate_inter_##SFX(const void* f, int64_t R, T* v, T* xs, T* xl, double tol,    \
                                     double step_size, int64_t steps, double zeta,                \
                                     uint8_t* assign, int64_t* steps_taken) {                     \
        return simulate_inter<T>(F(f), R, v, xs, xl, tol, step_size, steps, zeta, assign,         \
                                 steps_taken);                                                    \
    }                                                                                             \
    void dmm_init_v0_##SFX(uint64_t seed, int64_t replica, int64_t N, T* v) {                     \
        for (int64_t i = 0; i < N; ++i) v[i] = v0_value<T>(v0_bits(seed, uint64_t(replica), uint64_t(i))); \
    }                                                                                             \
    void dmm_batch_fixed_##SFX(const void* f, int64_t R, T* v, T* xs, T* xl, double dt,           \
                               double zeta, int64_t steps, int freeze, int64_t* solved_step,      \
                               int nthreads) {                                                    \
        batch_fixed<T>(F(f), R, v, xs, xl, T(dt), T(zeta), steps, freeze, solved_step, nthreads); \
    }                                                                                             \
    void dmm_batch_adaptive_##SFX(const void* f, int64_t R, T* v, T* xs, T* xl, double tol,       \
                                  double zeta, int64_t steps, int64_t* solved_step, T* dt_out,    \
                                  int nthreads) {                                                 \
        batch_adaptive<T>(F(f), R, v, xs, xl, T(tol), T(zeta), steps, solved_step, dt_out,        \
                          nthreads);                                                              \
    }

ORACLE_API(double, f64)
ORACLE_API(float, f32)

double dmm_default_zeta(const void* f) { return double(default_zeta<double>(F(f))); }
int dmm_evaluate_cnf(const void* f, const uint8_t* assign) { return evaluate_cnf(F(f), assign); }

int dmm_stoch_step(const void* f, uint8_t* v, uint64_t* xl, uint64_t seed, int64_t replica, int64_t step) {
    std::vector<uint64_t> a, b;
    return stoch_step(F(f), v, xl, seed, (uint64_t)replica, (uint64_t)step, a, b) ? 1 : 0;
}
void dmm_stoch_batch(const void* f, int64_t R, uint8_t* v, uint64_t* xl, uint64_t seed, int64_t replica_offset,
                     int64_t steps, int64_t* solved_step) {
    stoch_batch(F(f), R, v, xl, seed, replica_offset, steps, solved_step);
}

}  // extern "C"
