// kernels_gather.cuh — the GENERAL engine: replica-major state, per-variable gather.
//
// State layout in HBM: v[N][Rp], xs[M][Rp], xl[M][Rp] with the replica index fastest (Rp = R
// padded to 32 elements), so the threads of a warp read one row for 32 consecutive replicas as
// one coalesced 128/256-byte request.  With R = 1 the same kernels degenerate to one thread
// per row (consecutive threads → consecutive rows), which is the single-instance path.
//
// dv/dt is a deterministic GATHER, not a float-atomic scatter: the thread that owns (variable
// i, replica r) walks i's occurrence list — sorted by (clause, literal position), i.e. the
// order in which the reference's sequential loop reaches `dy.v[i] += …` (system.rs:35-80) —
// re-evaluates each clause's min / second-min from the read-only input state and adds the
// contributions starting from 0.0.  Sums are therefore bit-identical to the reference's.
// State is double-buffered (read t, write t+1) so one launch is one Euler step with no
// grid-wide barrier.
#pragma once
#include "common.cuh"

namespace odesat {

enum GatherMode { G_FIXED = 0, G_DERIV = 1, G_ADAPT_A = 2, G_ADAPT_B = 3 };

template <typename T> struct GatherArgs {
    FormulaDev f;
    int64_t R = 0, Rp = 0;
    const T *v = nullptr, *xs = nullptr, *xl = nullptr;   // state the RHS is evaluated on
    T *ov = nullptr, *oxs = nullptr, *oxl = nullptr;      // FIXED: y(t+1); DERIV: dy; A: y_half; B: y_new
    T *fv = nullptr, *fxs = nullptr, *fxl = nullptr;      // A: y_full (out); B: y_full (in)
    const T *yv = nullptr, *yxs = nullptr, *yxl = nullptr;   // B: the step's original y (copied through when done)
    T dt = T(0);
    const T* dt_arr = nullptr;    // per-replica dt (adaptive); overrides dt
    T zeta = T(0);
    T xl_max = T(0);              // 1e4 * M (system.rs:95)
    int32_t* solved_step = nullptr;   // [R] first flagged step, -1 = none
    uint32_t* unsat = nullptr;    // FIXED: ring [3][Rp]; others: [Rp]
    typename ErrBits<T>::U* err = nullptr;   // [Rp] (B)
    int32_t step = 0;
    int32_t freeze = 0;
};

// system.rs:43-57: running min / second-min over the literals of clause m for one replica.
template <typename T, int K>
__device__ __forceinline__ void clause_min2(const FormulaDev& f, const T* __restrict__ v, int64_t Rp,
                                            int64_t rep, int m, T& mn, T& sm) {
    mn = inf_v<T>();
    sm = inf_v<T>();
    int b, e;
    if (K > 0) { b = m * K; e = b + K; }
    else { b = __ldg(f.coff + m); e = __ldg(f.coff + m + 1); }
#pragma unroll
    for (int j = b; j < e; ++j) {
        const int lit = __ldg(f.lits + j);
        const int var = (lit < 0 ? -lit : lit) - 1;
        const T q = lit < 0 ? T(-1) : T(1);
        const T vi = __ldg(v + (int64_t)var * Rp + rep);
        const T val = T(1) - q * vi;                       // :49
        if (val < mn) { sm = mn; mn = val; }               // :50-52
        else if (val < sm) { sm = val; }                   // :53-55
    }
}

template <typename T, int K, int MODE>
__global__ void __launch_bounds__(256) k_gather(const GatherArgs<T> a) {
    const int64_t rep = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t row = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    const int64_t N = a.f.N, M = a.f.M, Rp = a.Rp;
    if (rep >= a.R || row >= N + M) return;

    bool passthru = false;   // replica is frozen/done: copy instead of integrate
    if (MODE == G_FIXED) {
        const int ss = a.solved_step[rep];
        const bool prev = (a.step > 0) && (a.unsat[(int64_t)((a.step + 2) % 3) * Rp + rep] == 0u);
        if (row == 0) {
            if (ss < 0 && prev) a.solved_step[rep] = a.step - 1;
            a.unsat[(int64_t)((a.step + 1) % 3) * Rp + rep] = 0u;
        }
        passthru = a.freeze && (ss >= 0 || prev);
    } else if (MODE == G_ADAPT_A) {
        if (a.solved_step[rep] >= 0) return;
    } else if (MODE == G_ADAPT_B) {
        passthru = (a.solved_step[rep] >= 0) || (a.unsat[rep] == 0u);
    }
    const T dt = a.dt_arr ? a.dt_arr[rep] : a.dt;

    if (row < N) {
        // ---------------- variable row: dv by ordered gather, then the v update ----------
        const int64_t at = row * Rp + rep;
        if (passthru) { a.ov[at] = (MODE == G_ADAPT_B) ? a.yv[at] : a.v[at]; return; }
        const T vi = __ldg(a.v + at);
        T dv = T(0);                                                     // :33
        const int e0 = __ldg(a.f.voff + row), e1 = __ldg(a.f.voff + row + 1);
        for (int e = e0; e < e1; ++e) {
            const int m = __ldg(a.f.occ_clause + e);
            T mn, sm;
            clause_min2<T, K>(a.f, a.v, Rp, rep, m, mn, sm);
            const T c = T(0.5) * mn;                                     // :60
            const int lit = __ldg(a.f.lits + __ldg(a.f.occ_slot + e));
            const T q = lit < 0 ? T(-1) : T(1);
            const T val = T(1) - q * vi;
            const T g = (T(0.5) * q) * ((val != mn) ? mn : sm);          // :64-70
            const T r = (c == val) ? T(0.5) * (q - vi) : T(0);           // :73-77
            const T xs_m = __ldg(a.xs + (int64_t)m * Rp + rep);
            const T xl_m = __ldg(a.xl + (int64_t)m * Rp + rep);
            dv = dv + ((xl_m * xs_m) * g + ((T(1) + a.zeta * xl_m) * (T(1) - xs_m)) * r);   // :80
        }
        if (MODE == G_DERIV) { a.ov[at] = dv; }
        else if (MODE == G_FIXED) { a.ov[at] = euler_clamp(vi, dv, dt, T(-1), T(1)); }   // :96
        else if (MODE == G_ADAPT_A) {
            a.ov[at] = euler_clamp(vi, dv, T(0.5) * dt, T(-1), T(1));    // :128
            a.fv[at] = euler_clamp(vi, dv, dt, T(-1), T(1));             // :125
        } else {
            const T yn = euler_clamp(vi, dv, T(0.5) * dt, T(-1), T(1));  // :130
            a.ov[at] = yn;
            const T e = fabs(a.fv[at] - yn);                             // :102-103
            if (e == e) {
                const typename ErrBits<T>::U b = ErrBits<T>::enc(e);
                if (b > a.err[rep]) atomicMax(a.err + rep, b);
            }
        }
    } else {
        // ---------------- clause row: C_m, the memories, the satisfied flag ----------------
        const int m = (int)(row - N);
        const int64_t at = (int64_t)m * Rp + rep;
        if (passthru) {
            a.oxs[at] = (MODE == G_ADAPT_B) ? a.yxs[at] : a.xs[at];
            a.oxl[at] = (MODE == G_ADAPT_B) ? a.yxl[at] : a.xl[at];
            return;
        }
        T mn, sm;
        clause_min2<T, K>(a.f, a.v, Rp, rep, m, mn, sm);
        const T c = T(0.5) * mn;
        const T xs_m = __ldg(a.xs + at), xl_m = __ldg(a.xl + at);
        const T dxs = (Kc<T>::BETA * (xs_m + Kc<T>::EPSILON)) * (c - Kc<T>::GAMMA);   // :84
        const T dxl = Kc<T>::ALPHA * (c - Kc<T>::DELTA);                               // :85
        const bool sat = c < Kc<T>::GAMMA;                                             // :88
        const T hi_s = T(1) - Kc<T>::EPSILON;
        if (MODE == G_FIXED) {
            if (!sat) a.unsat[(int64_t)(a.step % 3) * Rp + rep] = 1u;
            a.oxs[at] = euler_clamp(xs_m, dxs, dt, Kc<T>::EPSILON, hi_s);              // :94
            a.oxl[at] = euler_clamp(xl_m, dxl, dt, T(1), a.xl_max);                    // :95
        } else if (MODE == G_DERIV) {
            if (!sat) a.unsat[rep] = 1u;
            a.oxs[at] = dxs;
            a.oxl[at] = dxl;
        } else if (MODE == G_ADAPT_A) {
            if (!sat) a.unsat[rep] = 1u;
            const T h = T(0.5) * dt;
            a.oxs[at] = euler_clamp(xs_m, dxs, h, Kc<T>::EPSILON, hi_s);
            a.oxl[at] = euler_clamp(xl_m, dxl, h, T(1), a.xl_max);
            a.fxs[at] = euler_clamp(xs_m, dxs, dt, Kc<T>::EPSILON, hi_s);
            a.fxl[at] = euler_clamp(xl_m, dxl, dt, T(1), a.xl_max);
        } else {
            const T h = T(0.5) * dt;
            const T ns = euler_clamp(xs_m, dxs, h, Kc<T>::EPSILON, hi_s);
            const T nl = euler_clamp(xl_m, dxl, h, T(1), a.xl_max);
            a.oxs[at] = ns;
            a.oxl[at] = nl;
            const T e1 = fabs(a.fxs[at] - ns), e2 = fabs(a.fxl[at] - nl);
            T e = rmax(e1, e2);   // NaN-ignoring, like the reference's folds
            if (e == e) {
                const typename ErrBits<T>::U b = ErrBits<T>::enc(e);
                if (b > a.err[rep]) atomicMax(a.err + rep, b);
            }
        }
    }
}

// After the last FIXED step of a run: fold that step's flags into solved_step.
__global__ void k_fold_flags(int32_t* solved_step, const uint32_t* unsat_ring, int64_t R, int64_t Rp,
                             int32_t last_step) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R || last_step < 0) return;
    if (solved_step[r] < 0 && unsat_ring[(int64_t)(last_step % 3) * Rp + r] == 0u) solved_step[r] = last_step;
}

// Adaptive pass C (system.rs:132-135): per replica, commit the flag or update dt.
template <typename T>
__global__ void k_adapt_c(int32_t* solved_step, uint32_t* unsat, typename ErrBits<T>::U* err, T* dt, T tol,
                          int64_t R, int32_t step) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    if (solved_step[r] < 0) {
        if (unsat[r] == 0u) solved_step[r] = step;   // allsat: state untouched (system.rs:122)
        else {
            const T e = ErrBits<T>::dec(err[r]);
            dt[r] = rmax(rmin(dt[r] * sqrt(tol / e), T(1e3)), T(0.0078125));   // :133-135
        }
    }
    unsat[r] = 0u;
    err[r] = ErrBits<T>::NONE;
}

// system.rs:93-97 as a stand-alone elementwise kernel (odesat_update_state).
template <typename T>
__global__ void k_update_state(T* v, T* xs, T* xl, const T* dv, const T* dxs, const T* dxl, T dt, int64_t N,
                               int64_t M, int64_t R, int64_t Rp, T xl_max) {
    const int64_t rep = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t row = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    if (rep >= R || row >= N + M) return;
    if (row < N) {
        const int64_t at = row * Rp + rep;
        v[at] = euler_clamp(v[at], dv[at], dt, T(-1), T(1));
    } else {
        const int64_t at = (row - N) * Rp + rep;
        xs[at] = euler_clamp(xs[at], dxs[at], dt, Kc<T>::EPSILON, T(1) - Kc<T>::EPSILON);
        xl[at] = euler_clamp(xl[at], dxl[at], dt, T(1), xl_max);
    }
}

// system.rs:101-109 as a stand-alone kernel (odesat_max_error).
template <typename T>
__global__ void k_max_error(const T* av, const T* axs, const T* axl, const T* bv, const T* bxs, const T* bxl,
                            int64_t N, int64_t M, int64_t R, int64_t Rp, typename ErrBits<T>::U* err) {
    const int64_t rep = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t row = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    if (rep >= R || row >= N + M) return;
    T e;
    if (row < N) e = fabs(av[row * Rp + rep] - bv[row * Rp + rep]);
    else {
        const int64_t at = (row - N) * Rp + rep;
        e = rmax(fabs(axs[at] - bxs[at]), fabs(axl[at] - bxl[at]));
    }
    if (e == e) {
        const typename ErrBits<T>::U b = ErrBits<T>::enc(e);
        if (b > err[rep]) atomicMax(err + rep, b);
    }
}

template <typename T>
__global__ void k_err_decode(const typename ErrBits<T>::U* err, double* out, int64_t R) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < R) out[r] = (double)ErrBits<T>::dec(err[r]);
}

// main.rs:283-289 on the device: v0 from the counter-based generator, xs0 by clause polarity
// (system.rs:362-372), xl0 = 1.
template <typename T>
__global__ void k_init_state(T* v, T* xs, T* xl, const int8_t* xs0, int64_t N, int64_t M, int64_t R,
                             int64_t Rp, uint64_t seed, int64_t replica_offset, int gen_v, int gen_xs,
                             int gen_xl) {
    const int64_t rep = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t row = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    if (rep >= R || row >= N + M) return;
    if (row < N) {
        if (gen_v) v[row * Rp + rep] = v0_value<T>(v0_bits(v0_key(seed, (uint64_t)(replica_offset + rep)), (uint64_t)row));
    } else {
        const int64_t at = (row - N) * Rp + rep;
        if (gen_xs) xs[at] = (T)xs0[row - N];
        if (gen_xl) xl[at] = T(1);
    }
}

// cnf.rs:246-264 on the device for every replica: threshold v > 0 (system.rs:238) and
// evaluate each clause exactly; bad[rep] = 1 when some clause is falsified.
template <typename T>
__global__ void k_verify(const FormulaDev f, const T* v, int64_t R, int64_t Rp, uint32_t* bad) {
    const int64_t rep = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t m = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    if (rep >= R || m >= f.M) return;
    const int b = __ldg(f.coff + m), e = __ldg(f.coff + m + 1);
    bool sat = false;
    for (int j = b; j < e; ++j) {
        const int lit = __ldg(f.lits + j);
        const int var = (lit < 0 ? -lit : lit) - 1;
        const bool val = v[(int64_t)var * Rp + rep] > T(0);
        sat = sat || (lit < 0 ? !val : val);
    }
    if (!sat) bad[rep] = 1u;
}

template <typename T>
__global__ void k_assignment(const T* v, int64_t N, int64_t Rp, int64_t rep, uint8_t* out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) out[i] = v[i * Rp + rep] > T(0) ? 1 : 0;
}

// min over replicas of (solved_step << 32 | global replica index); INT64_MAX when none.
__global__ void k_first_key(const int32_t* solved_step, int64_t R, int64_t replica_offset,
                            unsigned long long* key) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int s = solved_step[r];
    if (s >= 0) atomicMin(key, ((unsigned long long)s << 32) | (unsigned long long)(replica_offset + r));
}

// Host layout [R][X] (one vector per replica, the reference's Vec<State>) ↔ replica-major [X][Rp].
template <typename T>
__global__ void k_transpose_in(const T* __restrict__ src, T* __restrict__ dst, int64_t R, int64_t X, int64_t Rp) {
    __shared__ T tile[32][33];
    const int64_t x0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int64_t r = r0 + k, x = x0 + threadIdx.x;
        if (r < R && x < X) tile[k][threadIdx.x] = src[r * X + x];
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int64_t x = x0 + k, r = r0 + threadIdx.x;
        if (r < R && x < X) dst[x * Rp + r] = tile[threadIdx.x][k];
    }
}
template <typename T>
__global__ void k_transpose_out(const T* __restrict__ src, T* __restrict__ dst, int64_t R, int64_t X, int64_t Rp) {
    __shared__ T tile[32][33];
    const int64_t x0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int64_t x = x0 + k, r = r0 + threadIdx.x;
        if (r < R && x < X) tile[k][threadIdx.x] = src[x * Rp + r];
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int64_t r = r0 + k, x = x0 + threadIdx.x;
        if (r < R && x < X) dst[r * X + x] = tile[threadIdx.x][k];
    }
}

}  // namespace odesat
