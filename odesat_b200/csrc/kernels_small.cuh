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
    const unsigned long long* stop_key = nullptr;   // see GatherArgs
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

// Shared-memory arrays of one resident replica.
template <typename T> struct SmallSmem {
    T *yv, *yxs, *yxl, *contrib, *hv, *hxs, *hxl, *fv, *fxs, *fxl;
    typename ErrBits<T>::U* red;
    __device__ SmallSmem(unsigned char* base, int N, int M, int L, bool adaptive) {
        yv = reinterpret_cast<T*>(base);
        yxs = yv + N;
        yxl = yxs + M;
        contrib = yxl + M;
        hv = contrib + L;          // adaptive only: y_half, y_full
        hxs = hv + N;
        hxl = hxs + M;
        fv = hxl + M;
        fxs = fv + N;
        fxl = fxs + M;
        red = reinterpret_cast<typename ErrBits<T>::U*>(
            base + (((adaptive ? 3 : 1) * (size_t)(N + 2 * M) + (size_t)L) * sizeof(T) + 15) / 16 * 16);
    }
};

// One step of the resident replica (all threads of the CTA; ends with the state complete and a
// block barrier passed).  Returns the all-clauses-satisfied flag of the PRE-update state (uniform).
//   fixed    (system.rs:141-154): the update always happens.
//   adaptive (system.rs:111-139): flagged ⇒ state and dt untouched; else one full step vs two half
//            steps, dt ← clamp(dt·sqrt(tol / err), 2^-7, 1e3).
template <typename T, int NT>
__device__ __forceinline__ bool small_step(const FormulaDev& f, const SmallSmem<T>& sm, bool adaptive, T& dt, T tol, T zeta, T xl_max) {
    using U = typename ErrBits<T>::U;
    const int N = (int)f.N, M = (int)f.M;
    const int tid = threadIdx.x;
    const T hi_s = T(1) - Kc<T>::EPSILON;
    // ---- k1 = f(y) ------------------------------------------------------------------------
    bool unsat = false;
    const T h = T(0.5) * dt;
    for (int m = tid; m < M; m += NT) {
        T dxs, dxl;
        const T x = sm.yxs[m], l = sm.yxl[m];
        unsat = !small_clause<T>(f, m, sm.yv, x, l, zeta, sm.contrib, dxs, dxl) || unsat;
        if (adaptive) {
            sm.hxs[m] = euler_clamp(x, dxs, h, Kc<T>::EPSILON, hi_s);      // :128
            sm.hxl[m] = euler_clamp(l, dxl, h, T(1), xl_max);
            sm.fxs[m] = euler_clamp(x, dxs, dt, Kc<T>::EPSILON, hi_s);     // :125
            sm.fxl[m] = euler_clamp(l, dxl, dt, T(1), xl_max);
        } else {
            sm.yxs[m] = euler_clamp(x, dxs, dt, Kc<T>::EPSILON, hi_s);     // :94 (own element only)
            sm.yxl[m] = euler_clamp(l, dxl, dt, T(1), xl_max);             // :95
        }
    }
    const bool allsat = !__syncthreads_or((int)unsat);            // :90 (also: contributions complete)
    if (!adaptive) {
        // :149-153 — the update happens even when the pre-update state was all-satisfied.  No thread
        // reads yv until the barrier below, so the in-place write is safe.
        for (int i = tid; i < N; i += NT) sm.yv[i] = euler_clamp(sm.yv[i], small_dv<T>(f, i, sm.contrib), dt, T(-1), T(1));   // :96
        __syncthreads();
        return allsat;
    }
    if (allsat) return true;                                       // :122 — state untouched
    for (int i = tid; i < N; i += NT) {
        const T dv = small_dv<T>(f, i, sm.contrib);
        sm.hv[i] = euler_clamp(sm.yv[i], dv, h, T(-1), T(1));      // :128
        sm.fv[i] = euler_clamp(sm.yv[i], dv, dt, T(-1), T(1));     // :125
    }
    __syncthreads();
    // ---- k2 = f(y_half); y_new = y_half + dt/2 k2; error vs y_full ---------------------------
    U err = ErrBits<T>::NONE;
    auto fold = [&](T e) { if (e == e) { const U b = ErrBits<T>::enc(e); if (b > err) err = b; } };   // NaN-ignoring max (:103)
    for (int m = tid; m < M; m += NT) {
        T dxs, dxl;
        const T x = sm.hxs[m], l = sm.hxl[m];
        small_clause<T>(f, m, sm.hv, x, l, zeta, sm.contrib, dxs, dxl);    // flag discarded (:129)
        const T nx = euler_clamp(x, dxs, h, Kc<T>::EPSILON, hi_s);         // :130
        const T nl = euler_clamp(l, dxl, h, T(1), xl_max);
        fold(fabs(sm.fxs[m] - nx));
        fold(fabs(sm.fxl[m] - nl));
        sm.yxs[m] = nx;
        sm.yxl[m] = nl;
    }
    __syncthreads();
    for (int i = tid; i < N; i += NT) {
        const T nv = euler_clamp(sm.hv[i], small_dv<T>(f, i, sm.contrib), h, T(-1), T(1));   // :130
        fold(fabs(sm.fv[i] - nv));
        sm.yv[i] = nv;
    }
    // block-wide NaN-ignoring max of the encoded errors
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const U other = __shfl_xor_sync(0xFFFFFFFFu, err, o); if (other > err) err = other; }
    if ((tid & 31) == 0) sm.red[tid >> 5] = err;
    __syncthreads();
    U tot = ErrBits<T>::NONE;
    for (int wq = 0; wq < NT / 32; ++wq) { const U o = sm.red[wq]; if (o > tot) tot = o; }
    const T e = ErrBits<T>::dec(tot);
    dt = rmax(rmin(dt * sqrt(tol / e), T(1e3)), T(0.0078125));     // :133-135
    __syncthreads();                                               // red is rewritten by the next step
    return false;
}

template <typename T, int NT>
__device__ __forceinline__ void small_load(const SmallArgs<T>& a, const SmallSmem<T>& sm, int64_t rep) {
    const int N = (int)a.f.N, M = (int)a.f.M;
    for (int i = threadIdx.x; i < N; i += NT) sm.yv[i] = a.v[(int64_t)i * a.Rp + rep];
    for (int m = threadIdx.x; m < M; m += NT) { sm.yxs[m] = a.xs[(int64_t)m * a.Rp + rep]; sm.yxl[m] = a.xl[(int64_t)m * a.Rp + rep]; }
}
template <typename T, int NT>
__device__ __forceinline__ void small_store(const SmallArgs<T>& a, const SmallSmem<T>& sm, int64_t rep) {
    const int N = (int)a.f.N, M = (int)a.f.M;
    for (int i = threadIdx.x; i < N; i += NT) a.v[(int64_t)i * a.Rp + rep] = sm.yv[i];
    for (int m = threadIdx.x; m < M; m += NT) { a.xs[(int64_t)m * a.Rp + rep] = sm.yxs[m]; a.xl[(int64_t)m * a.Rp + rep] = sm.yxl[m]; }
}

// `simulate` / `batch`: one CTA per replica, the whole chunk of steps with the replica resident.
template <typename T, int NT>
__global__ void __launch_bounds__(NT) k_solve_small(const SmallArgs<T> a) {
    if (a.stop_key != nullptr && *a.stop_key != 0x7FFFFFFFFFFFFFFFull) return;
    extern __shared__ __align__(16) unsigned char small_smem[];
    const SmallSmem<T> sm(small_smem, (int)a.f.N, (int)a.f.M, (int)a.f.L, a.adaptive != 0);
    const int64_t rep = blockIdx.x;
    small_load<T, NT>(a, sm, rep);
    int32_t solved_at = a.solved_step[rep];
    T dt = a.adaptive ? a.dt_arr[rep] : a.dt;
    __syncthreads();
    for (int s = 0; s < a.nsteps; ++s) {
        if (solved_at >= 0 && (a.adaptive || a.freeze)) break;     // flagged replicas are done (adaptive) / frozen (fixed)
        const bool allsat = small_step<T, NT>(a.f, sm, a.adaptive != 0, dt, a.tol, a.zeta, a.xl_max);
        if (allsat && solved_at < 0) solved_at = a.step0 + s;
    }
    small_store<T, NT>(a, sm, rep);
    if (threadIdx.x == 0) {
        a.solved_step[rep] = solved_at;
        if (a.adaptive) a.dt_arr[rep] = dt;
    }
}

// Adaptive `simulate_inter` (system.rs:312-349) EXACTLY as the reference runs it: the replicas take their
// steps one after the other inside every outer step and share ONE dt (`&mut dt`, system.rs:314 — quirk Q7),
// so replica r+1's step size is the one replica r's error estimate just produced.  That is a sequential
// dependency replica → replica: one CTA walks (outer step, replica) in order with the active replica
// resident in shared memory; the loop ends after the first outer step in which some replica flagged.
// dt_arr[0] carries the shared dt; solved_step[r] = the outer step on which r flagged.
template <typename T, int NT>
__global__ void __launch_bounds__(NT) k_inter_adaptive_small(const SmallArgs<T> a, int32_t* steps_done) {
    extern __shared__ __align__(16) unsigned char small_smem[];
    const SmallSmem<T> sm(small_smem, (int)a.f.N, (int)a.f.M, (int)a.f.L, true);
    T dt = a.dt_arr[0];
    int s = 0;
    bool any = false;
    for (; s < a.nsteps && !any; ++s) {
        for (int64_t rep = 0; rep < a.R; ++rep) {
            small_load<T, NT>(a, sm, rep);
            __syncthreads();
            const bool allsat = small_step<T, NT>(a.f, sm, true, dt, a.tol, a.zeta, a.xl_max);
            if (allsat) {
                any = true;
                if (threadIdx.x == 0 && a.solved_step[rep] < 0) a.solved_step[rep] = a.step0 + s;
            } else {
                small_store<T, NT>(a, sm, rep);
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        a.dt_arr[0] = dt;
        *steps_done = s;
    }
}

}  // namespace odesat
