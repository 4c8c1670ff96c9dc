// kernels_gather.cuh — the GENERAL engine: replica-major state, two phases per RHS evaluation.
//
// State layout in HBM: v[N][Rp], xs[M][Rp], xl[M][Rp] with the replica index fastest (Rp = R
// padded to 32 elements when R >= 32), so the threads of a warp read one row for consecutive
// replicas as one coalesced request (16 bytes = 4 f32 / 2 f64 replicas per thread).  With R = 1
// the same kernels degenerate to one thread per row (consecutive threads → consecutive rows),
// which is the single-instance path.
//
// dv/dt is a deterministic scatter-then-gather, not a float-atomic scatter:
//   clause phase   thread (clause m, replicas r..) evaluates the clause once — min / second-min,
//                  C_m, the memory derivatives, the satisfied flag — and writes the contribution
//                  of each of its literals to contrib[slot][r] (slot = position in the CSR);
//   variable phase thread (variable i, replicas r..) adds contrib[slot][r] over i's occurrence
//                  list, which is sorted by (clause, literal position) — the order in which the
//                  reference's sequential loop reaches `dy.v[i] += …` (system.rs:35-80) —
//                  starting from 0.0.  Sums are therefore bit-identical to the reference's.
// Every clause is evaluated exactly once per RHS (a single-pass gather would re-evaluate it once
// per literal: 15 state words per clause instead of 5).  FIXED steps update the state in place.
#pragma once
#include "common.cuh"
#include "packed_f32x2.cuh"

namespace odesat {

enum GatherMode { G_FIXED = 0, G_DERIV = 1, G_ADAPT_A = 2, G_ADAPT_B = 3 };

template <typename T> struct GatherArgs {
    FormulaDev f;
    int64_t R = 0, Rp = 0;
    int64_t rep0 = 0, rep1 = 0;   // replicas [rep0, rep1) of this launch (a SLAB; the whole batch when not slabbed)
    int64_t cstride = 0;          // row stride of contrib (Rp, or the slab width: contrib[slot][rep - rep0])
    int rows_per_thread = 8;      // rows a thread walks per launch: fewer, fatter blocks
    int l2_hints = 0;             // evict-first on the once-per-step streams, evict-last on the v gathers
    int fast = 0;                 // streaming clause kernel: every v is finite in [-1, 1] and zeta is finite, so the
                                  // tile kernel's shorter arithmetic (bit-identical on that domain) may be used
    const T *v = nullptr, *xs = nullptr, *xl = nullptr;   // state the RHS is evaluated on
    T *ov = nullptr, *oxs = nullptr, *oxl = nullptr;      // FIXED: y(t+1) (may alias the inputs); DERIV: dy; A: y_half; B: y_new
    T *fv = nullptr, *fxs = nullptr, *fxl = nullptr;      // A: y_full (out); B: y_full (in)
    const T *yv = nullptr, *yxs = nullptr, *yxl = nullptr;   // B: the step's original y (copied through when done)
    T* contrib = nullptr;         // [L][cstride] per-literal contributions to dv
    T dt = T(0);
    const T* dt_arr = nullptr;    // per-replica dt (adaptive); overrides dt
    T zeta = T(0);
    T xl_max = T(0);              // 1e4 * M (system.rs:95)
    int32_t* solved_step = nullptr;   // [R] first flagged step, -1 = none
    uint32_t* unsat = nullptr;    // [Rp] "some clause of this replica is unsatisfied" for the current RHS
    typename ErrBits<T>::U* err = nullptr;   // [Rp] (B)
    int32_t step = 0;
    int32_t freeze = 0;
    // `inter` early exit without a host round trip: the (all-reduced) key of the previous chunk of steps; when it
    // names a flagged replica, the speculatively issued launches of this chunk do nothing
    const unsigned long long* stop_key = nullptr;
};

constexpr unsigned long long GATHER_KEY_NONE = 0x7FFFFFFFFFFFFFFFull;
template <typename T> __device__ __forceinline__ bool gather_stopped(const GatherArgs<T>& a) {
    return a.stop_key != nullptr && *a.stop_key != GATHER_KEY_NONE;
}

template <typename T> __device__ __forceinline__ void err_max(typename ErrBits<T>::U* slot, T e) {
    if (e == e) {   // NaN-ignoring, like the reference's folds (system.rs:103)
        const typename ErrBits<T>::U b = ErrBits<T>::enc(e);
        if (b > *slot) atomicMax(slot, b);
    }
}

// V consecutive replicas per thread (16-byte accesses when V·sizeof(T) == 16, scalar when V == 1).
template <typename T, int V> struct RVec { T x[V]; };
template <typename T, int V> __device__ __forceinline__ RVec<T, V> vload(const T* p) {   // read-only in this launch
    RVec<T, V> r;
    if (V * sizeof(T) == 16) {
        const int4 q = __ldg(reinterpret_cast<const int4*>(p));
        *reinterpret_cast<int4*>(r.x) = q;
    } else {
#pragma unroll
        for (int u = 0; u < V; ++u) r.x[u] = __ldg(p + u);
    }
    return r;
}
template <typename T, int V> __device__ __forceinline__ RVec<T, V> vload_rw(const T* p) {   // data this launch may also write
    RVec<T, V> r;
    if (V * sizeof(T) == 16) *reinterpret_cast<int4*>(r.x) = *reinterpret_cast<const int4*>(p);
    else {
#pragma unroll
        for (int u = 0; u < V; ++u) r.x[u] = p[u];
    }
    return r;
}
template <typename T, int V> __device__ __forceinline__ void vstore(T* p, const RVec<T, V>& r) {
    if (V * sizeof(T) == 16) *reinterpret_cast<int4*>(p) = *reinterpret_cast<const int4*>(r.x);
    else {
#pragma unroll
        for (int u = 0; u < V; ++u) p[u] = r.x[u];
    }
}


// streaming (evict-first) variants for data that is touched once per step: keeps the L2 for the v rows,
// which every clause of a variable re-reads
template <typename T, int V> __device__ __forceinline__ void vstore_cs(T* p, const RVec<T, V>& r) {
    if (V * sizeof(T) == 16) __stcs(reinterpret_cast<int4*>(p), *reinterpret_cast<const int4*>(r.x));
    else {
#pragma unroll
        for (int u = 0; u < V; ++u) __stcs(p + u, r.x[u]);
    }
}
template <typename T, int V> __device__ __forceinline__ RVec<T, V> vload_cs(const T* p) {
    RVec<T, V> r;
    if (V * sizeof(T) == 16) *reinterpret_cast<int4*>(r.x) = __ldcs(reinterpret_cast<const int4*>(p));
    else {
#pragma unroll
        for (int u = 0; u < V; ++u) r.x[u] = __ldcs(p + u);
    }
    return r;
}

// ---- clause phase: system.rs:41-88 for one clause and V replicas ----------------------------
template <typename T, int K, int MODE, int V>
__device__ __forceinline__ void clause_row(const GatherArgs<T>& a, int64_t m, int64_t rep) {
    const int64_t Rp = a.Rp;
    const int64_t at = m * Rp + rep;
    bool skip[V];   // element is frozen / done / beyond R: keep (or pass through) its state
    bool all_skip = true;
#pragma unroll
    for (int u = 0; u < V; ++u) {
        const bool in = rep + u < a.R;
        bool sk = !in;
        if (in) {
            if (MODE == G_FIXED) sk = a.freeze && a.solved_step[rep + u] >= 0;
            else if (MODE == G_ADAPT_A) sk = a.solved_step[rep + u] >= 0;
            else if (MODE == G_ADAPT_B) sk = a.solved_step[rep + u] >= 0 || a.unsat[rep + u] == 0u;
        }
        skip[u] = sk;
        all_skip = all_skip && sk;
    }
    if (all_skip) {
        if (MODE == G_ADAPT_B) {   // done / just flagged: state untouched (system.rs:122)
            vstore<T, V>(a.oxs + at, vload<T, V>(a.yxs + at));
            vstore<T, V>(a.oxl + at, vload<T, V>(a.yxl + at));
        }
        return;
    }
    T dt[V];
#pragma unroll
    for (int u = 0; u < V; ++u) dt[u] = a.dt_arr ? a.dt_arr[rep + u < a.R ? rep + u : rep] : a.dt;
    int b, e;
    if (K > 0) { b = (int)m * K; e = b + K; }
    else { b = __ldg(a.f.coff + m); e = __ldg(a.f.coff + m + 1); }
    T mn[V], sm[V];
#pragma unroll
    for (int u = 0; u < V; ++u) { mn[u] = inf_v<T>(); sm[u] = inf_v<T>(); }
    RVec<T, V> vis[K > 0 ? K : 1];
    T qs[K > 0 ? K : 1];
#pragma unroll
    for (int j = b; j < e; ++j) {                              // :46-57
        const int lit = __ldg(a.f.lits + j);
        const int var = (lit < 0 ? -lit : lit) - 1;
        const T q = lit < 0 ? T(-1) : T(1);
        const RVec<T, V> vi = vload_rw<T, V>(a.v + (int64_t)var * Rp + rep);
#pragma unroll
        for (int u = 0; u < V; ++u) {
            const T val = T(1) - q * vi.x[u];                  // :49
            if (val < mn[u]) { sm[u] = mn[u]; mn[u] = val; }   // :50-52
            else if (val < sm[u]) { sm[u] = val; }             // :53-55
        }
        if (K > 0) { vis[j - b] = vi; qs[j - b] = q; }
    }
    const RVec<T, V> xs_m = vload_rw<T, V>(a.xs + at), xl_m = vload_rw<T, V>(a.xl + at);
    T c[V], w[V], rg[V];
#pragma unroll
    for (int u = 0; u < V; ++u) {
        c[u] = T(0.5) * mn[u];                                 // :60
        w[u] = xl_m.x[u] * xs_m.x[u];
        rg[u] = (T(1) + a.zeta * xl_m.x[u]) * (T(1) - xs_m.x[u]);
    }
#pragma unroll
    for (int j = b; j < e; ++j) {                              // :62-81
        RVec<T, V> vi;
        T q;
        if (K > 0) { vi = vis[j - b]; q = qs[j - b]; }
        else {
            const int lit = __ldg(a.f.lits + j);
            const int var = (lit < 0 ? -lit : lit) - 1;
            q = lit < 0 ? T(-1) : T(1);
            vi = vload_rw<T, V>(a.v + (int64_t)var * Rp + rep);
        }
        RVec<T, V> t;
#pragma unroll
        for (int u = 0; u < V; ++u) {
            const T val = T(1) - q * vi.x[u];
            const T g = (T(0.5) * q) * ((val != mn[u]) ? mn[u] : sm[u]);         // :64-70
            const T r = (c[u] == val) ? T(0.5) * (q - vi.x[u]) : T(0);           // :73-77
            t.x[u] = w[u] * g + rg[u] * r;                                       // the addend of :80
        }
        vstore<T, V>(a.contrib + (int64_t)j * a.cstride + (rep - a.rep0), t);
    }
    const T hi_s = T(1) - Kc<T>::EPSILON;
    RVec<T, V> o1, o2, f1, f2;
#pragma unroll
    for (int u = 0; u < V; ++u) {
        const T x = xs_m.x[u], l = xl_m.x[u];
        const T dxs = (Kc<T>::BETA * (x + Kc<T>::EPSILON)) * (c[u] - Kc<T>::GAMMA);   // :84
        const T dxl = Kc<T>::ALPHA * (c[u] - Kc<T>::DELTA);                            // :85
        // :88 (B's flag is discarded, :129); raised only, so a stale cached 1 just skips the store
        if (MODE != G_ADAPT_B && !skip[u] && !(c[u] < Kc<T>::GAMMA) && a.unsat[rep + u] == 0u) a.unsat[rep + u] = 1u;
        if (MODE == G_FIXED) {
            o1.x[u] = skip[u] ? x : euler_clamp(x, dxs, dt[u], Kc<T>::EPSILON, hi_s);  // :94
            o2.x[u] = skip[u] ? l : euler_clamp(l, dxl, dt[u], T(1), a.xl_max);        // :95
        } else if (MODE == G_DERIV) {
            o1.x[u] = dxs;
            o2.x[u] = dxl;
        } else if (MODE == G_ADAPT_A) {
            const T h = T(0.5) * dt[u];
            o1.x[u] = euler_clamp(x, dxs, h, Kc<T>::EPSILON, hi_s);                    // :128
            o2.x[u] = euler_clamp(l, dxl, h, T(1), a.xl_max);
            f1.x[u] = euler_clamp(x, dxs, dt[u], Kc<T>::EPSILON, hi_s);                // :125
            f2.x[u] = euler_clamp(l, dxl, dt[u], T(1), a.xl_max);
        } else {
            const T h = T(0.5) * dt[u];
            o1.x[u] = euler_clamp(x, dxs, h, Kc<T>::EPSILON, hi_s);                    // :130
            o2.x[u] = euler_clamp(l, dxl, h, T(1), a.xl_max);
        }
    }
    if (MODE == G_ADAPT_A) {
        // skipped elements of a mixed vector: their H/F rows are never read (pass B copies y through)
        vstore<T, V>(a.fxs + at, f1);
        vstore<T, V>(a.fxl + at, f2);
    }
    if (MODE == G_ADAPT_B) {
        const RVec<T, V> yx = vload<T, V>(a.yxs + at), yl = vload<T, V>(a.yxl + at);
        const RVec<T, V> fx = vload<T, V>(a.fxs + at), fl = vload<T, V>(a.fxl + at);
#pragma unroll
        for (int u = 0; u < V; ++u) {
            if (skip[u]) { o1.x[u] = yx.x[u]; o2.x[u] = yl.x[u]; }
            else if (rep + u < a.R) err_max<T>(a.err + rep + u, rmax(fabs(fx.x[u] - o1.x[u]), fabs(fl.x[u] - o2.x[u])));   // :104-107
        }
    }
    vstore<T, V>(a.oxs + at, o1);
    vstore<T, V>(a.oxl + at, o2);
}

template <typename T, int K, int MODE, int V>
__global__ void __launch_bounds__(256) k_clause_phase(const GatherArgs<T> a) {
    if (gather_stopped(a)) return;
    const int64_t rep = a.rep0 + ((int64_t)blockIdx.y * blockDim.x + threadIdx.x) * V;
    if (rep >= a.rep1) return;
#pragma unroll 2
    for (int it = 0; it < a.rows_per_thread; ++it) {
        const int64_t m = ((int64_t)blockIdx.x * a.rows_per_thread + it) * blockDim.y + threadIdx.y;
        if (m < a.f.M) clause_row<T, K, MODE, V>(a, m, rep);
    }
}

__device__ __forceinline__ float gfma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double gfma(double a, double b, double c) { return __fma_rn(a, b, c); }

// any |v| > 1 or NaN, or a memory outside mem_in_fast_domain, in the batch?  (decides whether the first step may use
// the fast arithmetic)
template <typename T>
__global__ void k_check_range(const T* v, const T* xs, const T* xl, int64_t N, int64_t M, int64_t R, int64_t Rp, unsigned* flag) {
    const int64_t rep = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t row = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    if (rep >= R || row >= N + M) return;
    if (row < N) {
        if (!(fabs(v[row * Rp + rep]) <= T(1))) *flag = 1u;
    } else {
        const int64_t m = row - N;
        if (!mem_in_fast_domain(xs[m * Rp + rep]) || !mem_in_fast_domain(xl[m * Rp + rep])) *flag = 1u;
    }
}

// ---- streaming clause phase (uniform 3-literal clauses) -------------------------------------------
// The plain kernel above is latency-bound: literal load → v gather → memory load are three
// dependent round trips per row (ncu: long-scoreboard stalls 16–17 cycles per issued instruction,
// 24 % of DRAM peak).  Here a thread owns RPT rows; it first loads the literals of all of them, then
// issues every state load of all of them — 3 v rows, xs, xl per row — as cp.async copies into its own
// shared-memory cells (no destination registers, so the loads of all RPT rows are in flight
// together), and only then computes row by row as the copy groups land.  Two round trips per RPT rows
// instead of three per row.  Same arithmetic, statement for statement, as clause_row.
template <int BYTES> __device__ __forceinline__ void cp_async_n(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
    else if (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
template <int BYTES> __device__ __forceinline__ void cp_async_n_hint(void* smem, const void* gmem, unsigned long long pol) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    if (BYTES == 16) asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "l"(pol) : "memory");
    else if (BYTES == 8) asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "l"(pol) : "memory");
    else asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;" ::"r"(s), "l"(gmem), "l"(pol) : "memory");
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void cp_async_commit_g() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_dyn(int n) {   // wait until at most n groups are pending
    switch (n) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    }
}

template <typename T, int MODE, int V, int RPT>
__global__ void __launch_bounds__(256) k_clause_stream(const GatherArgs<T> a) {
    static_assert(RPT >= 1 && RPT <= 4, "cp_async_wait_dyn covers up to 4 groups");
    if (gather_stopped(a)) return;
    constexpr int CB = V * (int)sizeof(T);                       // bytes of one cell (V replicas of one row)
    extern __shared__ __align__(16) unsigned char stream_smem[];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    auto cell = [&](int r, int c) { return stream_smem + ((size_t)(r * 5 + c) * 256 + tid) * CB; };
    const int64_t Rp = a.Rp;
    const int64_t rep = a.rep0 + ((int64_t)blockIdx.y * blockDim.x + threadIdx.x) * V;
    const bool rep_ok = rep < a.rep1;

    bool skip[V];   // element is frozen / done / beyond R: keep (or pass through) its state
    bool all_skip = true;
    T dt[V];
#pragma unroll
    for (int u = 0; u < V; ++u) {
        const bool in = rep_ok && rep + u < a.R;
        bool sk = !in;
        if (in) {
            if (MODE == G_FIXED) sk = a.freeze && a.solved_step[rep + u] >= 0;
            else if (MODE == G_ADAPT_A) sk = a.solved_step[rep + u] >= 0;
            else if (MODE == G_ADAPT_B) sk = a.solved_step[rep + u] >= 0 || a.unsat[rep + u] == 0u;
        }
        skip[u] = sk;
        all_skip = all_skip && sk;
        dt[u] = (a.dt_arr && rep_ok) ? a.dt_arr[rep + u < a.R ? rep + u : rep] : a.dt;
    }
    const unsigned long long pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    int lit[RPT][3];
    int64_t mrow[RPT];
    bool ok[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        mrow[r] = ((int64_t)blockIdx.x * RPT + r) * blockDim.y + threadIdx.y;
        ok[r] = rep_ok && mrow[r] < a.f.M;
        if (ok[r] && !all_skip) {
#pragma unroll
            for (int j = 0; j < 3; ++j) lit[r][j] = __ldg(a.f.lits + mrow[r] * 3 + j);
        }
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        if (ok[r] && !all_skip) {
            const int64_t at = mrow[r] * Rp + rep;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int var = (lit[r][j] < 0 ? -lit[r][j] : lit[r][j]) - 1;
                if (a.l2_hints) cp_async_n_hint<CB>(cell(r, j), a.v + (int64_t)var * Rp + rep, pol_keep);
                else cp_async_n<CB>(cell(r, j), a.v + (int64_t)var * Rp + rep);
            }
            if (a.l2_hints) {
                cp_async_n_hint<CB>(cell(r, 3), a.xs + at, pol_stream);
                cp_async_n_hint<CB>(cell(r, 4), a.xl + at, pol_stream);
            } else {
                cp_async_n<CB>(cell(r, 3), a.xs + at);
                cp_async_n<CB>(cell(r, 4), a.xl + at);
            }
        }
        cp_async_commit_g();
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        cp_async_wait_dyn(RPT - 1 - r);
        if (!ok[r]) continue;
        const int64_t at = mrow[r] * Rp + rep;
        if (all_skip) {
            if (MODE == G_ADAPT_B) {   // done / just flagged: state untouched (system.rs:122)
                vstore<T, V>(a.oxs + at, vload<T, V>(a.yxs + at));
                vstore<T, V>(a.oxl + at, vload<T, V>(a.yxl + at));
            }
            continue;
        }
        RVec<T, V> vis[3], xs_m, xl_m;
        T qs[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            vis[j] = *reinterpret_cast<const RVec<T, V>*>(cell(r, j));
            qs[j] = lit[r][j] < 0 ? T(-1) : T(1);
        }
        xs_m = *reinterpret_cast<const RVec<T, V>*>(cell(r, 3));
        xl_m = *reinterpret_cast<const RVec<T, V>*>(cell(r, 4));
        if constexpr (MODE == G_FIXED && sizeof(T) == 4 && V == 4) {
            // Fixed step, f32, four replicas per thread, fast domain: the tile kernel's arithmetic for two PAIRS of
            // replicas (packed f32x2: half the instructions of the scalar fast path below; the kernel was issue-bound
            // once the v gathers hit L2).  Same operations per lane, so the same bits.  A skipped replica (frozen, or
            // padding beyond R) integrates with dt = 0: its memories are already inside their clamp ranges.
            if (a.fast == 1) {
                const float q3[3] = {qs[0], qs[1], qs[2]};
                RVec<T, V> t[3], o1, o2;
                float mxv[4];
#pragma unroll
                for (int k = 0; k < 4; k += 2) {
                    const float2 v2[3] = {make_float2(vis[0].x[k], vis[0].x[k + 1]), make_float2(vis[1].x[k], vis[1].x[k + 1]),
                                          make_float2(vis[2].x[k], vis[2].x[k + 1])};
                    float2 d2[3] = {make_float2(0.0f, 0.0f), make_float2(0.0f, 0.0f), make_float2(0.0f, 0.0f)};
                    float2 xs2 = make_float2(xs_m.x[k], xs_m.x[k + 1]), xl2 = make_float2(xl_m.x[k], xl_m.x[k + 1]);
                    float mxk[2] = {0.0f, 0.0f};
                    const float2 dt2 = make_float2(skip[k] ? 0.0f : dt[k], skip[k + 1] ? 0.0f : dt[k + 1]);
                    clause_math_f32x2(v2, d2, q3, xs2, xl2, mxk, dt2, a.xl_max);
#pragma unroll
                    for (int j = 0; j < 3; ++j) { t[j].x[k] = d2[j].x; t[j].x[k + 1] = d2[j].y; }
                    o1.x[k] = skip[k] ? xs_m.x[k] : xs2.x; o1.x[k + 1] = skip[k + 1] ? xs_m.x[k + 1] : xs2.y;
                    o2.x[k] = skip[k] ? xl_m.x[k] : xl2.x; o2.x[k + 1] = skip[k + 1] ? xl_m.x[k + 1] : xl2.y;
                    mxv[k] = mxk[0]; mxv[k + 1] = mxk[1];
                }
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    if (a.l2_hints) vstore_cs<T, V>(a.contrib + (mrow[r] * 3 + j) * a.cstride + (rep - a.rep0), t[j]);
                    else vstore<T, V>(a.contrib + (mrow[r] * 3 + j) * a.cstride + (rep - a.rep0), t[j]);
                }
                // system.rs:88 — C_m >= 0.25 ⇔ min >= 0.5 exactly (C_m = 0.5·min, a power-of-two scaling)
#pragma unroll
                for (int u = 0; u < V; ++u)
                    if (!skip[u] && !(mxv[u] < 0.5f) && a.unsat[rep + u] == 0u) a.unsat[rep + u] = 1u;
                if (a.l2_hints) { vstore_cs<T, V>(a.oxs + at, o1); vstore_cs<T, V>(a.oxl + at, o2); }
                else { vstore<T, V>(a.oxs + at, o1); vstore<T, V>(a.oxl + at, o2); }
                continue;
            }
        }
        T mn[V], sm[V], c[V];
        if (a.fast) {
            // Fast arithmetic (see clause_math in tile_engine.cuh): 1 − q·v as one exact FMA, min / second-min as a
            // 5-op network, q·((0.5·xl·xs)·sel) for the addend, rigidity term dropped (identically ±0, quirk Q1).
            // Bit-identical sums on the domain v ∈ [-1, 1] finite, zeta finite.
            T av[3][V];
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int u = 0; u < V; ++u) av[j][u] = gfma(-qs[j], vis[j].x[u], T(1));
#pragma unroll
            for (int u = 0; u < V; ++u) {
                const T lo = rmin(av[0][u], av[1][u]), hi = rmax(av[0][u], av[1][u]);
                mn[u] = rmin(lo, av[2][u]);
                sm[u] = rmax(lo, rmin(hi, av[2][u]));
                c[u] = T(0.5) * mn[u];                             // :60
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                RVec<T, V> t;
#pragma unroll
                for (int u = 0; u < V; ++u) {
                    const T h = T(0.5) * (xl_m.x[u] * xs_m.x[u]);
                    t.x[u] = gfma(h * ((av[j][u] != mn[u]) ? mn[u] : sm[u]), qs[j], T(0));   // :64-70, :80
                }
                if (a.l2_hints) vstore_cs<T, V>(a.contrib + (mrow[r] * 3 + j) * a.cstride + (rep - a.rep0), t);
                else vstore<T, V>(a.contrib + (mrow[r] * 3 + j) * a.cstride + (rep - a.rep0), t);
            }
        } else {
#pragma unroll
            for (int u = 0; u < V; ++u) { mn[u] = inf_v<T>(); sm[u] = inf_v<T>(); }
#pragma unroll
            for (int j = 0; j < 3; ++j) {                              // :46-57
#pragma unroll
                for (int u = 0; u < V; ++u) {
                    const T val = T(1) - qs[j] * vis[j].x[u];          // :49
                    if (val < mn[u]) { sm[u] = mn[u]; mn[u] = val; }   // :50-52
                    else if (val < sm[u]) { sm[u] = val; }             // :53-55
                }
            }
            T w[V], rg[V];
#pragma unroll
            for (int u = 0; u < V; ++u) {
                c[u] = T(0.5) * mn[u];                                 // :60
                w[u] = xl_m.x[u] * xs_m.x[u];
                rg[u] = (T(1) + a.zeta * xl_m.x[u]) * (T(1) - xs_m.x[u]);
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) {                              // :62-81
                RVec<T, V> t;
#pragma unroll
                for (int u = 0; u < V; ++u) {
                    const T val = T(1) - qs[j] * vis[j].x[u];
                    const T g = (T(0.5) * qs[j]) * ((val != mn[u]) ? mn[u] : sm[u]);       // :64-70
                    const T rr = (c[u] == val) ? T(0.5) * (qs[j] - vis[j].x[u]) : T(0);    // :73-77
                    t.x[u] = w[u] * g + rg[u] * rr;                                        // the addend of :80
                }
                if (a.l2_hints) vstore_cs<T, V>(a.contrib + (mrow[r] * 3 + j) * a.cstride + (rep - a.rep0), t);
                else vstore<T, V>(a.contrib + (mrow[r] * 3 + j) * a.cstride + (rep - a.rep0), t);
            }
        }
        const T hi_s = T(1) - Kc<T>::EPSILON;
        RVec<T, V> o1, o2, f1, f2;
#pragma unroll
        for (int u = 0; u < V; ++u) {
            const T x = xs_m.x[u], l = xl_m.x[u];
            const T dxs = (Kc<T>::BETA * (x + Kc<T>::EPSILON)) * (c[u] - Kc<T>::GAMMA);   // :84
            const T dxl = Kc<T>::ALPHA * (c[u] - Kc<T>::DELTA);                            // :85
            // :88 (B's flag is discarded, :129); the flag is only ever raised, so a stale cached 1 just skips the store
            if (MODE != G_ADAPT_B && !skip[u] && !(c[u] < Kc<T>::GAMMA) && a.unsat[rep + u] == 0u) a.unsat[rep + u] = 1u;
            if (MODE == G_FIXED) {
                o1.x[u] = skip[u] ? x : euler_clamp(x, dxs, dt[u], Kc<T>::EPSILON, hi_s);  // :94
                o2.x[u] = skip[u] ? l : euler_clamp(l, dxl, dt[u], T(1), a.xl_max);        // :95
            } else if (MODE == G_DERIV) {
                o1.x[u] = dxs;
                o2.x[u] = dxl;
            } else if (MODE == G_ADAPT_A) {
                const T h = T(0.5) * dt[u];
                o1.x[u] = euler_clamp(x, dxs, h, Kc<T>::EPSILON, hi_s);                    // :128
                o2.x[u] = euler_clamp(l, dxl, h, T(1), a.xl_max);
                f1.x[u] = euler_clamp(x, dxs, dt[u], Kc<T>::EPSILON, hi_s);                // :125
                f2.x[u] = euler_clamp(l, dxl, dt[u], T(1), a.xl_max);
            } else {
                const T h = T(0.5) * dt[u];
                o1.x[u] = euler_clamp(x, dxs, h, Kc<T>::EPSILON, hi_s);                    // :130
                o2.x[u] = euler_clamp(l, dxl, h, T(1), a.xl_max);
            }
        }
        if (MODE == G_ADAPT_A) {
            vstore<T, V>(a.fxs + at, f1);
            vstore<T, V>(a.fxl + at, f2);
        }
        if (MODE == G_ADAPT_B) {
            const RVec<T, V> yx = vload<T, V>(a.yxs + at), yl = vload<T, V>(a.yxl + at);
            const RVec<T, V> fx = vload<T, V>(a.fxs + at), fl = vload<T, V>(a.fxl + at);
#pragma unroll
            for (int u = 0; u < V; ++u) {
                if (skip[u]) { o1.x[u] = yx.x[u]; o2.x[u] = yl.x[u]; }
                else if (rep + u < a.R) err_max<T>(a.err + rep + u, rmax(fabs(fx.x[u] - o1.x[u]), fabs(fl.x[u] - o2.x[u])));   // :104-107
            }
        }
        if (a.l2_hints) { vstore_cs<T, V>(a.oxs + at, o1); vstore_cs<T, V>(a.oxl + at, o2); }
        else { vstore<T, V>(a.oxs + at, o1); vstore<T, V>(a.oxl + at, o2); }
    }
}

// ---- variable phase: the ordered sum of :80 and the v update of :96 ----------------------------
// Row 0 of every replica also commits the replica's flag (FIXED): the clause phase of step s has
// finished when this kernel runs, so unsat[r] is the complete AND of :90.
template <typename T, int MODE, int V>
__device__ __forceinline__ void var_row(const GatherArgs<T>& a, int64_t row, int64_t rep) {
    const int64_t N = a.f.N, Rp = a.Rp;
    bool skip[V];
    bool all_skip = true;
#pragma unroll
    for (int u = 0; u < V; ++u) {
        const bool in = rep + u < a.R;
        bool sk = !in;
        if (in) {
            if (MODE == G_FIXED) {
                const int ss = a.solved_step[rep + u];
                sk = a.freeze && ss >= 0 && ss < a.step;      // flagged in an EARLIER step
                if (row == 0) {
                    if (ss < 0 && a.unsat[rep + u] == 0u) a.solved_step[rep + u] = a.step;   // the update below still happens (:149-153)
                    a.unsat[rep + u] = 0u;
                }
            } else if (MODE == G_ADAPT_A) sk = a.solved_step[rep + u] >= 0;
            else if (MODE == G_ADAPT_B) sk = a.solved_step[rep + u] >= 0 || a.unsat[rep + u] == 0u;
        }
        skip[u] = sk;
        all_skip = all_skip && sk;
    }
    if (row >= N) return;
    const int64_t at = row * Rp + rep;
    if (all_skip) {
        if (MODE == G_ADAPT_B) vstore<T, V>(a.ov + at, vload<T, V>(a.yv + at));
        return;
    }
    const RVec<T, V> vi = vload_rw<T, V>(a.v + at);
    T dv[V];
#pragma unroll
    for (int u = 0; u < V; ++u) dv[u] = T(0);                         // :33
    const int e0 = __ldg(a.f.voff + row), e1 = __ldg(a.f.voff + row + 1);
    for (int e = e0; e < e1; e += 4) {                                // 4 independent row loads in flight, added in order
        RVec<T, V> t[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (e + k < e1) {
                const T* cp = a.contrib + (int64_t)__ldg(a.f.occ_slot + e + k) * a.cstride + (rep - a.rep0);
                t[k] = a.l2_hints ? vload_cs<T, V>(cp) : vload_rw<T, V>(cp);
            }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (e + k < e1) {
#pragma unroll
                for (int u = 0; u < V; ++u) dv[u] = dv[u] + t[k].x[u];   // :80
            }
    }
    RVec<T, V> o, fo;
    if (MODE == G_ADAPT_B) {
        const RVec<T, V> y = vload<T, V>(a.yv + at), fv = vload<T, V>(a.fv + at);
#pragma unroll
        for (int u = 0; u < V; ++u) {
            const T dt = a.dt_arr ? a.dt_arr[rep + u < a.R ? rep + u : rep] : a.dt;
            const T yn = euler_clamp(vi.x[u], dv[u], T(0.5) * dt, T(-1), T(1));      // :130
            if (skip[u]) o.x[u] = y.x[u];
            else {
                o.x[u] = yn;
                if (rep + u < a.R) err_max<T>(a.err + rep + u, fabs(fv.x[u] - yn));   // :102-103
            }
        }
        vstore<T, V>(a.ov + at, o);
        return;
    }
#pragma unroll
    for (int u = 0; u < V; ++u) {
        const T dt = a.dt_arr ? a.dt_arr[rep + u < a.R ? rep + u : rep] : a.dt;
        if (MODE == G_DERIV) o.x[u] = dv[u];
        else if (MODE == G_FIXED) o.x[u] = skip[u] ? vi.x[u] : euler_clamp(vi.x[u], dv[u], dt, T(-1), T(1));   // :96
        else {
            o.x[u] = euler_clamp(vi.x[u], dv[u], T(0.5) * dt, T(-1), T(1));          // :128
            fo.x[u] = euler_clamp(vi.x[u], dv[u], dt, T(-1), T(1));                  // :125
        }
    }
    vstore<T, V>(a.ov + at, o);
    if (MODE == G_ADAPT_A) vstore<T, V>(a.fv + at, fo);
}

template <typename T, int MODE, int V>
__global__ void __launch_bounds__(256) k_var_phase(const GatherArgs<T> a) {
    if (gather_stopped(a)) return;
    const int64_t rep = a.rep0 + ((int64_t)blockIdx.y * blockDim.x + threadIdx.x) * V;
    if (rep >= a.rep1) return;
    const int64_t rows = a.f.N > 0 ? a.f.N : 1;
    for (int it = 0; it < a.rows_per_thread; ++it) {
        const int64_t row = ((int64_t)blockIdx.x * a.rows_per_thread + it) * blockDim.y + threadIdx.y;
        if (row < rows) var_row<T, MODE, V>(a, row, rep);
    }
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
    if (!sat && bad[rep] == 0u) bad[rep] = 1u;   // raised only: test first, every falsified clause of a replica hits the same word
}

template <typename T>
__global__ void k_assignment(const T* v, int64_t N, int64_t Rp, int64_t rep, uint8_t* out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) out[i] = v[i * Rp + rep] > T(0) ? 1 : 0;
}

// min over replicas of (solved_step << 32 | global replica index), INT64_MAX when none, and the number of replicas
// that have not flagged — ONE block, plain stores (no reset launch): out[0] = key, out[1] = unflagged count.
// The early-exit word of `inter` / the all-done test of `batch`; 8 + 8 bytes per chunk of steps.
__global__ void __launch_bounds__(1024) k_first_key(const int32_t* solved_step, int64_t R, int64_t replica_offset,
                                                    unsigned long long* out) {
    __shared__ unsigned long long s_key[32];
    __shared__ unsigned s_cnt[32];
    unsigned long long key = 0x7FFFFFFFFFFFFFFFull;
    unsigned cnt = 0;
    for (int64_t r = threadIdx.x; r < R; r += blockDim.x) {
        const int s = solved_step[r];
        if (s >= 0) {
            const unsigned long long k = ((unsigned long long)s << 32) | (unsigned long long)(replica_offset + r);
            key = k < key ? k : key;
        } else {
            ++cnt;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long k2 = __shfl_xor_sync(0xFFFFFFFFu, key, o);
        key = k2 < key ? k2 : key;
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0) { s_key[threadIdx.x >> 5] = key; s_cnt[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int nw = (blockDim.x + 31) >> 5;
        key = (int)threadIdx.x < nw ? s_key[threadIdx.x] : 0x7FFFFFFFFFFFFFFFull;
        cnt = (int)threadIdx.x < nw ? s_cnt[threadIdx.x] : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long k2 = __shfl_xor_sync(0xFFFFFFFFu, key, o);
            key = k2 < key ? k2 : key;
            cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
        }
        if (threadIdx.x == 0) { out[0] = key; out[1] = cnt; }
    }
}

// dt of every replica back to the initial step size (system.rs:205) without a host buffer
template <typename T> __global__ void k_fill(T* p, int64_t n, T x) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = x;
}

// Host layout [R][X] (one vector per replica, the reference's Vec<State>) ↔ replica-major [X][Rp].
// rows of a single instance between the caller's clause order and the sorted view's storage order (formula.hpp):
// gather: dst[p] = src[perm[p]];  scatter: dst[perm[p]] = src[p]
template <typename T>
__global__ void k_permute_rows(const T* __restrict__ src, T* __restrict__ dst, const int32_t* __restrict__ perm, int64_t X, int gather) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= X) return;
    if (gather) dst[p] = src[perm[p]];
    else dst[perm[p]] = src[p];
}

template <typename T>
// perm (may be NULL): device row x holds the caller's row perm[x] (sorted clause view, formula.hpp)
__global__ void k_transpose_in(const T* __restrict__ src, T* __restrict__ dst, int64_t R, int64_t X, int64_t Rp,
                               const int32_t* __restrict__ perm) {
    __shared__ T tile[32][33];
    const int64_t x0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int64_t r = r0 + k, x = x0 + threadIdx.x;
        if (r < R && x < X) tile[k][threadIdx.x] = src[r * X + (perm ? (int64_t)perm[x] : x)];
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int64_t x = x0 + k, r = r0 + threadIdx.x;
        if (r < R && x < X) dst[x * Rp + r] = tile[threadIdx.x][k];
    }
}
template <typename T>
__global__ void k_transpose_out(const T* __restrict__ src, T* __restrict__ dst, int64_t R, int64_t X, int64_t Rp,
                                const int32_t* __restrict__ perm) {
    __shared__ T tile[32][33];
    const int64_t x0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int64_t x = x0 + k, r = r0 + threadIdx.x;
        if (r < R && x < X) tile[k][threadIdx.x] = src[x * Rp + r];
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int64_t r = r0 + k, x = x0 + threadIdx.x;
        if (r < R && x < X) dst[r * X + (perm ? (int64_t)perm[x] : x)] = tile[threadIdx.x][k];
    }
}

}  // namespace odesat
