// kernels_small.cuh — persistent small-instance integrator (any clause lengths, fixed AND
// adaptive steps): ONE CTA owns one replica for a whole chunk of steps, the complete state — y,
// y_half, y_full and the per-literal contributions — lives in shared memory, and the step loop of
// `simulate` (system.rs:190-233) runs inside the kernel: no launch, no HBM traffic per step.
//
// This is the path of `solve` / adaptive `batch` on the reference's own fixtures (aim-100: N = 100,
// M = 160; after `-r 7` preprocessing ≈ 60 variables and ragged clauses): the general engine needs
// five launches per adaptive step there and is launch-bound (≈ 20 µs per step); here a step is a few
// block barriers (≈ 1 µs).  The arithmetic is the reference's statement for statement (same
// expressions as clause_row / var_row of kernels_gather.cuh), dv summed in the reference's order.
#pragma once
#include "common.cuh"

namespace odesat {

template <typename T> struct SmallArgs {
    FormulaDev f;
    int64_t R = 0, Rp = 0;
    T *v = nullptr, *xs = nullptr, *xl = nullptr;   // canonical [row][Rp], updated in place
    T* dt_arr = nullptr;                              // [R] adaptive step size (in/out)
    int32_t* solved_step = nullptr;                   // [R] first flagged step, -1 = none
    T dt = T(0), tol = T(0), zeta = T(0), xl_max = T(0);
    int32_t step0 = 0, nsteps = 0, freeze = 0, adaptive = 0;
};

// elements of T the kernel keeps in shared memory
inline size_t small_smem_bytes(int64_t N, int64_t M, int64_t L, bool adaptive, size_t elem) {
    const size_t state = (size_t)(N + 2 * M);
    return ((adaptive ? 3 : 1) * state + (size_t)L) * elem + 64 * 8 + 64;
}

// system.rs:41-88 for one clause: contributions of its literals into contrib[slot], the memory
// derivatives, and the clause-satisfied test.
template <typename T>
__device__ __forceinline__ bool small_clause(const FormulaDev& f, int m, const T* __restrict__ v, T xs_m, T xl_m, T zeta,
                                             T* __restrict__ contrib, T& dxs, T& dxl) {
    const int b = __ldg(f.coff + m), e = __ldg(f.coff + m + 1);
    T mn = inf_v<T>(), sm = inf_v<T>();
    for (int j = b; j < e; ++j) {                                  // :46-57
        const int lit = __ldg(f.lits + j);
        const int var = (lit < 0 ? -lit : lit) - 1;
        const T q = lit < 0 ? T(-1) : T(1);
        const T val = T(1) - q * v[var];                           // :49
        if (val < mn) { sm = mn; mn = val; }                       // :50-52
        else if (val < sm) { sm = val; }                           // :53-55
    }
    const T c = T(0.5) * mn;                                       // :60
    const T w = xl_m * xs_m;
    const T rg = (T(1) + zeta * xl_m) * (T(1) - xs_m);
    for (int j = b; j < e; ++j) {                                  // :62-81
        const int lit = __ldg(f.lits + j);
        const int var = (lit < 0 ? -lit : lit) - 1;
        const T q = lit < 0 ? T(-1) : T(1);
        const T vi = v[var];
        const T val = T(1) - q * vi;
        const T g = (T(0.5) * q) * ((val != mn) ? mn : sm);        // :64-70
        const T r = (c == val) ? T(0.5) * (q - vi) : T(0);         // :73-77
        contrib[j] = w * g + rg * r;                               // the addend of :80
    }
    dxs = (Kc<T>::BETA * (xs_m + Kc<T>::EPSILON)) * (c - Kc<T>::GAMMA);   // :84
    dxl = Kc<T>::ALPHA * (c - Kc<T>::DELTA);                              // :85
    return c < Kc<T>::GAMMA;                                              // :88
}

template <typename T> __device__ __forceinline__ T small_dv(const FormulaDev& f, int i, const T* __restrict__ contrib) {
    T dv = T(0);                                                   // :33
    const int e0 = __ldg(f.voff + i), e1 = __ldg(f.voff + i + 1);
    for (int e = e0; e < e1; ++e) dv = dv + contrib[__ldg(f.occ_slot + e)];   // :80, reference order
    return dv;
}

template <typename T, int NT>
__global__ void __launch_bounds__(NT) k_solve_small(const SmallArgs<T> a) {
    using U = typename ErrBits<T>::U;
    extern __shared__ __align__(16) unsigned char small_smem[];
    const int N = (int)a.f.N, M = (int)a.f.M, L = (int)a.f.L;
    const int tid = threadIdx.x;
    const int64_t rep = blockIdx.x;
    T* yv = reinterpret_cast<T*>(small_smem);
    T* yxs = yv + N;
    T* yxl = yxs + M;
    T* contrib = yxl + M;
    T* hv = contrib + L;          // adaptive only: y_half, y_full
    T* hxs = hv + N;
    T* hxl = hxs + M;
    T* fv = hxl + M;
    T* fxs = fv + N;
    T* fxl = fxs + M;
    U* s_red = reinterpret_cast<U*>(small_smem + (((a.adaptive ? 3 : 1) * (size_t)(N + 2 * M) + (size_t)L) * sizeof(T) + 15) / 16 * 16);

    const int64_t Rp = a.Rp;
    for (int i = tid; i < N; i += NT) yv[i] = a.v[(int64_t)i * Rp + rep];
    for (int m = tid; m < M; m += NT) { yxs[m] = a.xs[(int64_t)m * Rp + rep]; yxl[m] = a.xl[(int64_t)m * Rp + rep]; }
    int32_t solved_at = a.solved_step[rep];
    T dt = a.adaptive ? a.dt_arr[rep] : a.dt;
    const T hi_s = T(1) - Kc<T>::EPSILON;
    __syncthreads();

    for (int s = 0; s < a.nsteps; ++s) {
        if (solved_at >= 0 && (a.adaptive || a.freeze)) break;     // flagged replicas are done (adaptive) / frozen (fixed)
        // ---- k1 = f(y) ------------------------------------------------------------------------
        bool unsat = false;
        const T h = T(0.5) * dt;
        for (int m = tid; m < M; m += NT) {
            T dxs, dxl;
            const T x = yxs[m], l = yxl[m];
            unsat = !small_clause<T>(a.f, m, yv, x, l, a.zeta, contrib, dxs, dxl) || unsat;
            if (a.adaptive) {
                hxs[m] = euler_clamp(x, dxs, h, Kc<T>::EPSILON, hi_s);      // :128
                hxl[m] = euler_clamp(l, dxl, h, T(1), a.xl_max);
                fxs[m] = euler_clamp(x, dxs, dt, Kc<T>::EPSILON, hi_s);     // :125
                fxl[m] = euler_clamp(l, dxl, dt, T(1), a.xl_max);
            } else {
                yxs[m] = euler_clamp(x, dxs, dt, Kc<T>::EPSILON, hi_s);     // :94 (own element only)
                yxl[m] = euler_clamp(l, dxl, dt, T(1), a.xl_max);           // :95
            }
        }
        const bool allsat = !__syncthreads_or((int)unsat);         // :90 (also: contributions complete)
        if (allsat && solved_at < 0) {
            solved_at = a.step0 + s;
            if (a.adaptive) break;                                 // :122 — state untouched
        }
        if (!a.adaptive) {
            // :149-153 — the update happens even when the pre-update state was all-satisfied.  No thread
            // reads yv until the barrier below, so the in-place write is safe.
            for (int i = tid; i < N; i += NT) yv[i] = euler_clamp(yv[i], small_dv<T>(a.f, i, contrib), dt, T(-1), T(1));   // :96
            __syncthreads();
            continue;
        }
        for (int i = tid; i < N; i += NT) {
            const T dv = small_dv<T>(a.f, i, contrib);
            hv[i] = euler_clamp(yv[i], dv, h, T(-1), T(1));        // :128
            fv[i] = euler_clamp(yv[i], dv, dt, T(-1), T(1));       // :125
        }
        __syncthreads();
        // ---- k2 = f(y_half); y_new = y_half + dt/2 k2; error vs y_full ---------------------------
        U err = ErrBits<T>::NONE;
        auto fold = [&](T e) { if (e == e) { const U b = ErrBits<T>::enc(e); if (b > err) err = b; } };   // NaN-ignoring max (:103)
        for (int m = tid; m < M; m += NT) {
            T dxs, dxl;
            const T x = hxs[m], l = hxl[m];
            small_clause<T>(a.f, m, hv, x, l, a.zeta, contrib, dxs, dxl);   // flag discarded (:129)
            const T nx = euler_clamp(x, dxs, h, Kc<T>::EPSILON, hi_s);      // :130
            const T nl = euler_clamp(l, dxl, h, T(1), a.xl_max);
            fold(fabs(fxs[m] - nx));
            fold(fabs(fxl[m] - nl));
            yxs[m] = nx;
            yxl[m] = nl;
        }
        __syncthreads();
        for (int i = tid; i < N; i += NT) {
            const T nv = euler_clamp(hv[i], small_dv<T>(a.f, i, contrib), h, T(-1), T(1));   // :130
            fold(fabs(fv[i] - nv));
            yv[i] = nv;
        }
        // block-wide NaN-ignoring max of the encoded errors
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const U other = __shfl_xor_sync(0xFFFFFFFFu, err, o); if (other > err) err = other; }
        if ((tid & 31) == 0) s_red[tid >> 5] = err;
        __syncthreads();
        U tot = ErrBits<T>::NONE;
        for (int wq = 0; wq < NT / 32; ++wq) { const U o = s_red[wq]; if (o > tot) tot = o; }
        const T e = ErrBits<T>::dec(tot);
        dt = rmax(rmin(dt * sqrt(a.tol / e), T(1e3)), T(0.0078125));   // :133-135
        __syncthreads();                                               // s_red is rewritten next step
    }

    for (int i = tid; i < N; i += NT) a.v[(int64_t)i * Rp + rep] = yv[i];
    for (int m = tid; m < M; m += NT) { a.xs[(int64_t)m * Rp + rep] = yxs[m]; a.xl[(int64_t)m * Rp + rep] = yxl[m]; }
    if (tid == 0) {
        a.solved_step[rep] = solved_at;
        if (a.adaptive) a.dt_arr[rep] = dt;
    }
}

}  // namespace odesat
