// system.hpp — C++ mirror of the reference's `odesat::system` module (src/system.rs) over the C ABI.
//
// The reference is compiled code (Rust) and this image has no Rust toolchain, so the host side
// above include/odesat_b200.h is mirrored in C++: same function names, argument order and
// meaning as the Rust `pub fn`s; `Option<T>` becomes std::optional<T>; panics become
// odesat::system::Error.  The `&mut SlabState` scratch argument has no counterpart (the kernels
// keep min / second-min in registers).  Header-only; link with -lodesat_b200.
#pragma once
#include <cmath>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/odesat_b200.h"

namespace odesat {

// cnf.rs:5-9, 30-33, 53-57
struct Literal { std::size_t variable; bool is_negated; };
struct CNFClause { std::vector<Literal> literals; };
struct CNFFormula { std::vector<CNFClause> clauses; std::size_t varnum = 0; };

namespace system {

struct Error : std::runtime_error {
    int code;
    Error(int c, const char* m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) { if (rc != ODESAT_OK) throw Error(rc, odesat_last_error()); }

// system.rs:6-11
struct State {
    std::vector<double> v, xs, xl;
};

// `&CNFFormula` flattened once to the CSR the ABI takes; owns the device copy.
class Formula {
public:
    explicit Formula(const CNFFormula& f) : varnum_(f.varnum), n_clauses_(f.clauses.size()) {
        std::vector<int64_t> off(1, 0);
        std::vector<int32_t> lits;
        for (const auto& c : f.clauses) {
            for (const auto& l : c.literals) lits.push_back(l.is_negated ? -int32_t(l.variable + 1) : int32_t(l.variable + 1));
            off.push_back((int64_t)lits.size());
        }
        check(odesat_formula_create((int64_t)f.varnum, (int64_t)f.clauses.size(), off.data(), lits.data(), &h_));
    }
    ~Formula() { odesat_formula_destroy(h_); }
    Formula(const Formula&) = delete;
    Formula& operator=(const Formula&) = delete;
    const odesat_formula* handle() const { return h_; }
    std::size_t varnum() const { return varnum_; }
    std::size_t n_clauses() const { return n_clauses_; }
private:
    odesat_formula* h_ = nullptr;
    std::size_t varnum_, n_clauses_;
};

inline odesat_params make_params(std::optional<double> tolerance, std::optional<double> step_size,
                                 std::optional<std::size_t> steps, std::optional<double> learning_rate) {
    odesat_params p{};
    p.tolerance = tolerance.value_or(NAN);
    p.step_size = step_size.value_or(NAN);
    p.steps = steps ? (int64_t)*steps : -1;
    p.learning_rate = learning_rate.value_or(NAN);
    p.precision = ODESAT_F64;
    p.engine = ODESAT_ENGINE_AUTO;
    p.schedule = ODESAT_SCHED_EXACT;
    p.chunk = 0;
    return p;
}

// system.rs:362-372
inline std::vector<double> init_short_term_memory(const Formula& f) {
    std::vector<double> xs(f.n_clauses());
    check(odesat_init_short_term_memory(f.handle(), xs.data()));
    return xs;
}
// system.rs:25-31
inline bool compute_derivatives(const State& y, State& dy, const Formula& f, double zeta) {
    int a = 0;
    dy.v.resize(y.v.size()); dy.xs.resize(y.xs.size()); dy.xl.resize(y.xl.size());
    check(odesat_compute_derivatives(f.handle(), y.v.data(), y.xs.data(), y.xl.data(), zeta, dy.v.data(), dy.xs.data(), dy.xl.data(), &a));
    return a != 0;
}
// system.rs:93
inline void update_state(State& s, const State& d, double dt, const Formula& f) {
    check(odesat_update_state(f.handle(), s.v.data(), s.xs.data(), s.xl.data(), d.v.data(), d.xs.data(), d.xl.data(), dt));
}
// system.rs:101
inline double max_error(const State& a, const State& b, const Formula& f) {
    double e = 0;
    check(odesat_max_error(f.handle(), a.v.data(), a.xs.data(), a.xl.data(), b.v.data(), b.xs.data(), b.xl.data(), &e));
    return e;
}
// system.rs:141-148
inline bool euler_step_fixed(State& s, const Formula& f, double dt, double zeta) {
    int a = 0;
    check(odesat_euler_step_fixed(f.handle(), s.v.data(), s.xs.data(), s.xl.data(), dt, zeta, &a));
    return a != 0;
}
// system.rs:111-119
inline bool euler_step(State& s, const Formula& f, double tolerance, double& dt, double zeta) {
    int a = 0;
    check(odesat_euler_step(f.handle(), s.v.data(), s.xs.data(), s.xl.data(), tolerance, &dt, zeta, &a));
    return a != 0;
}
// system.rs:156-163
inline std::vector<bool> simulate(State& s, const Formula& f, std::optional<double> tolerance, std::optional<double> step_size,
                                  std::optional<std::size_t> steps, std::optional<double> learning_rate) {
    const odesat_params p = make_params(tolerance, step_size, steps, learning_rate);
    std::vector<uint8_t> a(f.varnum());
    check(odesat_simulate(f.handle(), s.v.data(), s.xs.data(), s.xl.data(), &p, a.data(), nullptr, nullptr, nullptr));
    return std::vector<bool>(a.begin(), a.end());
}
// system.rs:241-248
inline std::vector<bool> simulate_inter(std::vector<State>& states, const Formula& f, std::optional<double> tolerance,
                                        std::optional<double> step_size, std::optional<std::size_t> steps,
                                        std::optional<double> learning_rate) {
    const std::size_t R = states.size(), N = f.varnum(), M = f.n_clauses();
    std::vector<double> v(R * N), xs(R * M), xl(R * M);
    for (std::size_t r = 0; r < R; ++r) {
        std::copy(states[r].v.begin(), states[r].v.end(), v.begin() + r * N);
        std::copy(states[r].xs.begin(), states[r].xs.end(), xs.begin() + r * M);
        std::copy(states[r].xl.begin(), states[r].xl.end(), xl.begin() + r * M);
    }
    const odesat_params p = make_params(tolerance, step_size, steps, learning_rate);
    std::vector<uint8_t> a(N);
    int64_t winner = -1, taken = 0;
    check(odesat_simulate_inter(f.handle(), (int64_t)R, v.data(), xs.data(), xl.data(), &p, a.data(), &winner, &taken));
    for (std::size_t r = 0; r < R; ++r) {
        states[r].v.assign(v.begin() + r * N, v.begin() + (r + 1) * N);
        states[r].xs.assign(xs.begin() + r * M, xs.begin() + (r + 1) * M);
        states[r].xl.assign(xl.begin() + r * M, xl.begin() + (r + 1) * M);
    }
    return std::vector<bool>(a.begin(), a.end());
}

}  // namespace system
}  // namespace odesat
