// tile_adaptive.cuh — ADAPTIVE Euler steps (system.rs:111-139) in the tile engine (included by tile_engine.cuh).
//
// `batch` without -s and `solve` integrate with the step-doubling controller of euler_step: per step two RHS
// evaluations (k1 on y, k2 on y_half), three updates, the ∞-norm of y_full − y_new and a new dt, per replica
// (main.rs:278-308 runs `simulate` once per replica).  Round 1 ran this on the gather engine only: five launches per
// step, y_half / y_full materialised in HBM, contributions written and re-read.  Here one CTA owns a replica tile for a
// whole chunk of adaptive steps, exactly like k_tile_fixed: {v, dv} rows in shared memory, clauses streamed level by
// level in the order of the compiled schedule (tile_schedule.hpp), one thread = one clause slot, the per-thread
// cp.async ring.  A step is two walks over the schedule:
//
//   pass A   k1 = f(y):  reads {xs, xl} and gathers v, accumulates dv, raises the all-satisfied flag (the only RHS
//            whose flag the reference uses, :120) and stores C_m — ONE scalar per clause and replica — instead of
//            the half- and full-step memories: dxs and dxl are functions of (xs, xl, C_m) alone (:84-85), so pass B
//            recomputes xs_half, xl_half, xs_full, xl_full bit for bit from the untouched {xs, xl} and C_m.
//            Variable pass A: v_full → a global scratch row (read back by the same thread), rows ← {v_half, 0}.
//            A replica whose flag is up is frozen here: "state untouched" (:122) costs nothing because pass A has
//            not written any state yet.
//   pass B   k2 = f(y_half): rebuilds the half/full memories, gathers v_half, accumulates dv, writes
//            {xs_new, xl_new} in place and folds |y_full − y_new| of the memories into a per-thread maximum.
//            Variable pass B: v_new = clamp(v_half + (dt/2)·dv), |v_full − v_new|; the error is reduced per replica
//            (warp reduce + shared-memory integer max on the bit pattern: NaN-ignoring like the folds of max_error,
//            :101-109) and every thread computes the replica's next dt (:133-135).
//
// HBM traffic per clause, replica and adaptive step: {xs, xl} read twice and written once + C_m written and read =
// 8 scalars, against 12 for the materialising form of SURVEY §8(d) (read y, write y_half and y_full; read both, write
// y_new); v never leaves the SM inside a chunk (v_full: 2 scalars per VARIABLE).
//
// Arithmetic: the statements of system.rs in their order (STRICT) or the tile kernel's fast forms that are bit-identical
// on the domain the integrator lives in (clause_math in tile_engine.cuh) — so with the EXACT schedule every replica's
// trajectory, dt sequence and flag step equal the oracle's bit for bit (tests/test_gpu_tile_adaptive.py).
#pragma once

namespace odesat {

template <typename T> struct TileAdaptArgs {
    TileArgs<T> t;          // geometry, schedule, state, flags, zeta, xl_max, step0 / nsteps, stop_key, oor (dt unused)
    T* vfull = nullptr;     // [tiles][N][W]     v of the full step (scratch)
    T* cm = nullptr;        // [tiles][Mpad][W]  C_m of pass A (scratch)
    T* dt = nullptr;        // [R] per-replica step size, read at launch start, written back at its end
    T tol = T(0);
    const uint32_t* aux = nullptr;   // literal words of the loop clauses (RAGGED instantiations; tile_schedule.hpp)
};

// One RHS evaluation of a clause for one replica (system.rs:43-88) without the update: contributions added into d,
// C_m returned.  Same two code paths as clause_math.
template <typename T, bool STRICT>
__device__ __forceinline__ T clause_rhs(const T (&v)[3], T (&d)[3], const T (&q)[3], T xs, T xl, T zeta) {
    T a[3], mn, sm;
    if (STRICT) {
        mn = inf_v<T>();
        sm = inf_v<T>();
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            a[j] = T(1) - q[j] * v[j];                                      // :49
            if (a[j] < mn) { sm = mn; mn = a[j]; } else if (a[j] < sm) { sm = a[j]; }   // :50-55
        }
    } else {
#pragma unroll
        for (int j = 0; j < 3; ++j) a[j] = fma_exact(-q[j], v[j], T(1));
        const T lo = rmin(a[0], a[1]), hi = rmax(a[0], a[1]);
        mn = rmin(lo, a[2]);
        sm = rmax(lo, rmin(hi, a[2]));
    }
    const T cm = T(0.5) * mn;                                               // :60
    const T wgt = xl * xs;
    if (STRICT) {
        const T rg = (T(1) + zeta * xl) * (T(1) - xs);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const T g = (T(0.5) * q[j]) * ((a[j] != mn) ? mn : sm);         // :64-70
            const T r = (cm == a[j]) ? T(0.5) * (q[j] - v[j]) : T(0);       // :73-77
            d[j] = d[j] + (wgt * g + rg * r);                               // :80
        }
    } else {
        const T h = T(0.5) * wgt;
#pragma unroll
        for (int j = 0; j < 3; ++j) d[j] = fma_exact(h * ((a[j] != mn) ? mn : sm), q[j], d[j]);
    }
    return cm;
}

// a C_m cell: W scalars = 8 bytes
__device__ __forceinline__ uint2 pack_cm(const float (&c)[2]) { return make_uint2(__float_as_uint(c[0]), __float_as_uint(c[1])); }
__device__ __forceinline__ uint2 pack_cm(const double (&c)[1]) { return make_uint2((unsigned)__double2loint(c[0]), (unsigned)__double2hiint(c[0])); }
__device__ __forceinline__ void unpack_cm(const uint2 u, float (&c)[2]) { c[0] = __uint_as_float(u.x); c[1] = __uint_as_float(u.y); }
__device__ __forceinline__ void unpack_cm(const uint2 u, double (&c)[1]) { c[0] = __hiloint2double((int)u.y, (int)u.x); }

template <typename U> __device__ __forceinline__ U warp_max_bits(U x);
template <> __device__ __forceinline__ unsigned warp_max_bits<unsigned>(unsigned x) { return __reduce_max_sync(0xFFFFFFFFu, x); }
template <> __device__ __forceinline__ unsigned long long warp_max_bits<unsigned long long>(unsigned long long x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long y = __shfl_xor_sync(0xFFFFFFFFu, x, o);
        x = y > x ? y : x;
    }
    return x;
}

// ---- the two f32 replicas of a tile at once (packed f32x2, same operations in the same order as the scalar fast path) ----
// RHS of one clause (system.rs:43-80): contributions added into d, → the clause minimum (C_m = 0.5·min exactly)
__device__ __forceinline__ float2 clause_rhs_f32x2(const float2 (&v)[3], float2 (&d)[3], const float (&q)[3], float2 xs, float2 xl) {
    float2 a[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) a[j] = fma2(bc2(-q[j]), v[j], bc2(1.0f));
    float2 mn, sm;
    {
        const float lo = rmin(a[0].x, a[1].x), hi = rmax(a[0].x, a[1].x);
        mn.x = rmin(lo, a[2].x);
        sm.x = rmax(lo, rmin(hi, a[2].x));
    }
    {
        const float lo = rmin(a[0].y, a[1].y), hi = rmax(a[0].y, a[1].y);
        mn.y = rmin(lo, a[2].y);
        sm.y = rmax(lo, rmin(hi, a[2].y));
    }
    const float2 h = mul2(bc2(0.5f), mul2(xl, xs));
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float2 sel = make_float2((a[j].x != mn.x) ? mn.x : sm.x, (a[j].y != mn.y) ? mn.y : sm.y);
        d[j] = fma2(mul2(h, sel), bc2(q[j]), d[j]);                         // :64-70, :80
    }
    return mn;
}
// y + dt·dy clamped: the product packed, the addition scalar (ptxas contracts a packed mul + add into FFMA2 even with
// -fmad=false, which would round once instead of twice — see packed_f32x2.cuh)
__device__ __forceinline__ float2 euler_clamp_f32x2(float2 y, float2 dy, float2 dt, float lo, float hi) {
    const float2 p = mul2(dt, dy);
    return make_float2(rmin(rmax(__fadd_rn(y.x, p.x), lo), hi), rmin(rmax(__fadd_rn(y.y, p.y), lo), hi));
}
// memory derivatives (system.rs:84-85) from C_m
__device__ __forceinline__ void mem_derivs_f32x2(float2 xs, float2 cm, float2& dxs, float2& dxl) {
    dxs = mul2(mul2(bc2(Kc<float>::BETA), add2(xs, bc2(Kc<float>::EPSILON))), add2(cm, bc2(-Kc<float>::GAMMA)));
    dxl = mul2(bc2(Kc<float>::ALPHA), add2(cm, bc2(-Kc<float>::DELTA)));
}

// RHS of one LOOP clause (no literal, more than three, or a repeated variable — tile_ragged.cuh) for the W replicas of the
// tile: system.rs:43-81 statement by statement, contributions added into the dv halves of the rows, → C_m.
template <typename T, int W>
__device__ __forceinline__ void clause_loop_rhs(unsigned char* smem_raw, const uint32_t* __restrict__ lits, unsigned len, const T (&xs)[W],
                                                const T (&xl)[W], T zeta, T (&cm)[W]) {
    using Row = typename TileTraits<T>::Row;
    using IO = RowIO<T, W>;
    T mn[W], sm[W];
#pragma unroll
    for (int w = 0; w < W; ++w) { mn[w] = inf_v<T>(); sm[w] = inf_v<T>(); }
    const uint4* lits4 = reinterpret_cast<const uint4*>(lits);
    for (unsigned j0 = 0; j0 < len; j0 += 4) {
        const uint4 wa = __ldg(lits4 + (j0 >> 2));
        const uint32_t lws[4] = {wa.x, wa.y, wa.z, wa.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (j0 + u < len) {
                const Row* r = reinterpret_cast<const Row*>(smem_raw + (lws[u] & 0x3FFF0u));
                const T q = (lws[u] >> 31) ? T(-1) : T(1);
                T v[W], d[W];
                IO::unpack(*r, v, d);
#pragma unroll
                for (int w = 0; w < W; ++w) {
                    const T a = T(1) - q * v[w];                                                    // :49
                    if (a < mn[w]) { sm[w] = mn[w]; mn[w] = a; } else if (a < sm[w]) { sm[w] = a; }   // :50-55
                }
            }
        }
    }
    T wgt[W], rg[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
        cm[w] = T(0.5) * mn[w];                                                             // :60
        wgt[w] = xl[w] * xs[w];
        rg[w] = (T(1) + zeta * xl[w]) * (T(1) - xs[w]);
    }
    for (unsigned j0 = 0; j0 < len; j0 += 4) {
        const uint4 wa = __ldg(lits4 + (j0 >> 2));
        const uint32_t lws[4] = {wa.x, wa.y, wa.z, wa.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (j0 + u < len) {   // in literal order: a repeated variable sees its own earlier addend
                Row* r = reinterpret_cast<Row*>(smem_raw + (lws[u] & 0x3FFF0u));
                const T q = (lws[u] >> 31) ? T(-1) : T(1);
                T v[W], d[W];
                IO::unpack(*r, v, d);
#pragma unroll
                for (int w = 0; w < W; ++w) {
                    const T a = T(1) - q * v[w];
                    const T g = (T(0.5) * q) * ((a != mn[w]) ? mn[w] : sm[w]);             // :64-70
                    const T rr = (cm[w] == a) ? T(0.5) * (q - v[w]) : T(0);                 // :73-77
                    d[w] = d[w] + (wgt[w] * g + rg[w] * rr);                                // :80
                }
                IO::store_dv(r, d);
            }
        }
    }
}

// RAGGED: the schedule may hold one- and two-literal packed clauses (missing positions enter with +inf and are not
// written) and LOOP clauses; group clauses (EXACT schedules of formulas with 4..32-literal clauses) are not handled here —
// TileEngine::has_adaptive() is false for those.
// Shared memory: rows[N] (16 B) | ring_m[D][NT] (16 B) | ring_e[D][NT] (8 B) | ring_c[D][NT] (8 B) | items[n_items]
template <typename T, int NT, int D, bool STRICT, bool RAGGED = false>
__global__ void __launch_bounds__(NT, 1) k_tile_adaptive(const TileAdaptArgs<T> aa) {
    constexpr int W = TileTraits<T>::W;
    using Row = typename TileTraits<T>::Row;
    using Mem = typename TileTraits<T>::Mem;
    using IO = RowIO<T, W>;
    using EB = ErrBits<T>;
    using U = typename EB::U;
    static_assert(sizeof(T) * W == 8, "a C_m cell is 8 bytes");
    const TileArgs<T>& a = aa.t;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Row* rows = reinterpret_cast<Row*>(smem_raw);
    Mem* ring_m = reinterpret_cast<Mem*>(smem_raw + (size_t)a.N * sizeof(Row));
    uint2* ring_e = reinterpret_cast<uint2*>(ring_m + D * NT);
    uint2* ring_c = ring_e + D * NT;
    uint2* s_items = ring_c + D * NT;   // {slot base, count | last << 31}
    __shared__ U s_err[W];

    const int s_first = launch_first_step<STRICT>(a);   // block-uniform
    if (s_first >= a.nsteps) return;
    const unsigned tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    static_assert(D == 2 || D == 3, "the item table is allocated with two spare entries");
    const int n_items = (a.n_items + D - 1) / D * D;
    const uint2* my_entry = reinterpret_cast<const uint2*>(a.entry) + tid;      // + slot base
    Mem* my_mem = a.mem + tile * a.Mpad + tid;                                  // + slot base
    uint2* my_cm = reinterpret_cast<uint2*>(aa.cm) + tile * a.Mpad + tid;       // + slot base
    Mem* my_cell_m = ring_m + tid;                                              // + k·NT
    uint2* my_cell_e = ring_e + tid;
    uint2* my_cell_c = ring_c + tid;
    T* vt = a.vt + tile * a.N * W;
    T* vfull = aa.vfull + tile * a.N * W;

    // the item list is walked in whole rings: padded here with empty items to a multiple of D (≤ a.n_items + 2 entries)
    for (int i = tid; i < n_items; i += NT) {
        const uint32_t it = i < a.n_items ? a.items[i] : 0u;
        s_items[i] = make_uint2(it & 0xFFFFFu, ((it >> 20) & 0x7FFu) | (it & TILE_ITEM_LAST));
    }
    for (int i = tid; i < a.N; i += NT) {
        T v[W], dv[W];
#pragma unroll
        for (int w = 0; w < W; ++w) { v[w] = vt[(int64_t)i * W + w]; dv[w] = T(0); }
        rows[i] = IO::pack(v, dv);
    }
    if (tid < W) s_err[tid] = EB::NONE;
    bool valid[W], frozen[W];
    int32_t solved_at[W];
    T dtw[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
        valid[w] = tile * W + w < a.R;
        solved_at[w] = valid[w] ? a.solved[tile * W + w] : 0;
        frozen[w] = !valid[w] || solved_at[w] >= 0;      // a flagged replica is never evaluated again (system.rs:229-231)
        dtw[w] = valid[w] ? aa.dt[tile * W + w] : T(0.01);
    }
    __syncthreads();

    // ring: item i of pass p sits in stage i % D; pass B's cells carry C_m as well
    auto fetch = [&](int k, const uint2 it, bool with_cm) {
        if (tid < (it.y & 0x7FFFFFFFu)) {
            cp_async16(my_cell_m + k * NT, at16(my_mem, it.x));
            cp_async8(my_cell_e + k * NT, at8(my_entry, it.x));
            if (with_cm) cp_async8(my_cell_c + k * NT, at8(my_cm, it.x));
        }
        cp_async_commit();
    };
#pragma unroll
    for (int k = 0; k < D; ++k) fetch(k, s_items[k], false);

    const T hi_s = T(1) - Kc<T>::EPSILON;
    for (int s = s_first; s < a.nsteps; ++s) {
        bool all_frozen = true;
#pragma unroll
        for (int w = 0; w < W; ++w) all_frozen = all_frozen && frozen[w];
        if (all_frozen) break;
        bool unsat[W];
        float mx[2] = {0.0f, 0.0f};   // f32x2 fast path: running max of pass A's clause minima (→ unsat)
        U e_loc[W];
        T hw[W];
#pragma unroll
        for (int w = 0; w < W; ++w) { unsat[w] = false; e_loc[w] = EB::NONE; hw[w] = T(0.5) * dtw[w]; }   // :128

        for (int pass = 0; pass < 2; ++pass) {
            // ------------------------------ clause phase ---------------------------------------
            for (int base = 0; base < n_items; base += D) {
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const int i = base + k;
                    const uint2 it = s_items[i];
                    cp_async_wait<D - 1>();                     // this thread's cells of item i have landed
                    if (tid < (it.y & 0x7FFFFFFFu)) {
                        const Mem mm = my_cell_m[k * NT];
                        const uint2 e = my_cell_e[k * NT];
                        Row* const r0 = reinterpret_cast<Row*>(smem_raw + (e.x & 0x3FFF0u));
                        Row* const r1 = reinterpret_cast<Row*>(smem_raw + ((e.x >> 14) & 0x3FFF0u));
                        Row* const r2 = reinterpret_cast<Row*>(smem_raw + (e.y & 0x3FFF0u));
                        const T q[3] = {(e.y >> 24) & 1u ? T(-1) : T(1), (e.y >> 25) & 1u ? T(-1) : T(1), (e.y >> 26) & 1u ? T(-1) : T(1)};
                        T v[3][W], d[3][W], xs[W], xl[W];
                        IO::unpack_mem(mm, xs, xl);
                        const bool loopc = RAGGED && (e.y & TILE_ENTRY_LOOP) != 0u;
                        const bool no2 = RAGGED && (e.y & TILE_ENTRY_NO2) != 0u, no1 = RAGGED && (e.y & TILE_ENTRY_NO1) != 0u;
                        if (!loopc) {
                            IO::unpack(*r0, v[0], d[0]);
                            IO::unpack(*r1, v[1], d[1]);
                            IO::unpack(*r2, v[2], d[2]);
                            if constexpr (RAGGED) {   // a literal position that does not exist: value +inf (system.rs:46-47)
#pragma unroll
                                for (int w = 0; w < W; ++w) {
                                    if (no2) v[2][w] = -inf_v<T>();
                                    if (no1) v[1][w] = -inf_v<T>();
                                }
                            }
                        }
                        constexpr bool PACKED = !STRICT && W == 2 && sizeof(T) == 4;   // packed f32x2 arithmetic (not for loop clauses)
                        auto scalar_clause = [&]() {
                        if (pass == 0) {
                            T cm[W];
                            if (loopc) clause_loop_rhs<T, W>(smem_raw, aa.aux + e.x, e.y & 0xFFFFu, xs, xl, a.zeta, cm);
#pragma unroll
                            for (int w = 0; w < W; ++w) {
                                if (!loopc) {
                                    const T vv[3] = {v[0][w], v[1][w], v[2][w]};
                                    T dd[3] = {d[0][w], d[1][w], d[2][w]};
                                    cm[w] = clause_rhs<T, STRICT>(vv, dd, q, xs[w], xl[w], a.zeta);
                                    d[0][w] = dd[0]; d[1][w] = dd[1]; d[2][w] = dd[2];
                                }
                                unsat[w] = unsat[w] || !(cm[w] < Kc<T>::GAMMA);                      // :88
                            }
                            // plain store (not .cg): the cp.async.ca of pass B reads it back through this SM's L1
                            *at8(my_cm, it.x) = pack_cm(cm);
                        } else {
                            T cm1[W], xs_f[W], xl_f[W], xs_h[W], xl_h[W], cm2[W];
                            unpack_cm(my_cell_c[k * NT], cm1);
#pragma unroll
                            for (int w = 0; w < W; ++w) {
                                const T dxs1 = (Kc<T>::BETA * (xs[w] + Kc<T>::EPSILON)) * (cm1[w] - Kc<T>::GAMMA);   // :84
                                const T dxl1 = Kc<T>::ALPHA * (cm1[w] - Kc<T>::DELTA);                               // :85
                                xs_f[w] = euler_clamp(xs[w], dxs1, dtw[w], Kc<T>::EPSILON, hi_s);                    // :125
                                xl_f[w] = euler_clamp(xl[w], dxl1, dtw[w], T(1), a.xl_max);
                                xs_h[w] = euler_clamp(xs[w], dxs1, hw[w], Kc<T>::EPSILON, hi_s);                     // :128
                                xl_h[w] = euler_clamp(xl[w], dxl1, hw[w], T(1), a.xl_max);
                            }
                            if (loopc) clause_loop_rhs<T, W>(smem_raw, aa.aux + e.x, e.y & 0xFFFFu, xs_h, xl_h, a.zeta, cm2);   // :129
#pragma unroll
                            for (int w = 0; w < W; ++w) {
                                if (!loopc) {
                                    const T vv[3] = {v[0][w], v[1][w], v[2][w]};
                                    T dd[3] = {d[0][w], d[1][w], d[2][w]};
                                    cm2[w] = clause_rhs<T, STRICT>(vv, dd, q, xs_h[w], xl_h[w], a.zeta);             // :129
                                    d[0][w] = dd[0]; d[1][w] = dd[1]; d[2][w] = dd[2];
                                }
                                const T dxs2 = (Kc<T>::BETA * (xs_h[w] + Kc<T>::EPSILON)) * (cm2[w] - Kc<T>::GAMMA);
                                const T dxl2 = Kc<T>::ALPHA * (cm2[w] - Kc<T>::DELTA);
                                const T xs_n = euler_clamp(xs_h[w], dxs2, hw[w], Kc<T>::EPSILON, hi_s);              // :130
                                const T xl_n = euler_clamp(xl_h[w], dxl2, hw[w], T(1), a.xl_max);
                                if (!frozen[w]) {
                                    const T e1 = fabs(xs_f[w] - xs_n), e2 = fabs(xl_f[w] - xl_n);                   // :104-107
                                    if (e1 == e1) { const U b = EB::enc(e1); e_loc[w] = b > e_loc[w] ? b : e_loc[w]; }
                                    if (e2 == e2) { const U b = EB::enc(e2); e_loc[w] = b > e_loc[w] ? b : e_loc[w]; }
                                    xs[w] = xs_n;
                                    xl[w] = xl_n;
                                }
                            }
                            __stcg(at16(my_mem, it.x), IO::pack_mem(xs, xl));
                        }
                        };
                        if constexpr (!PACKED) scalar_clause();
                        else if (loopc) scalar_clause();
                        else {
                            const float2 v2[3] = {make_float2(v[0][0], v[0][1]), make_float2(v[1][0], v[1][1]), make_float2(v[2][0], v[2][1])};
                            float2 d2[3] = {make_float2(d[0][0], d[0][1]), make_float2(d[1][0], d[1][1]), make_float2(d[2][0], d[2][1])};
                            const float qf[3] = {(float)q[0], (float)q[1], (float)q[2]};
                            const float2 xs2 = make_float2(xs[0], xs[1]), xl2 = make_float2(xl[0], xl[1]);
                            if (pass == 0) {
                                const float2 mn = clause_rhs_f32x2(v2, d2, qf, xs2, xl2);
                                mx[0] = rmax(mx[0], mn.x);              // :88 as a running maximum of the clause minima (packed_f32x2.cuh)
                                mx[1] = rmax(mx[1], mn.y);
                                const float2 cm = mul2(bc2(0.5f), mn);                                                   // :60
                                *at8(my_cm, it.x) = make_uint2(__float_as_uint(cm.x), __float_as_uint(cm.y));
                            } else {
                                const uint2 cu = my_cell_c[k * NT];
                                const float2 cm1 = make_float2(__uint_as_float(cu.x), __uint_as_float(cu.y));
                                const float2 dt2 = make_float2(dtw[0], dtw[1]), h2 = make_float2(hw[0], hw[1]);
                                float2 dxs1, dxl1, dxs2, dxl2;
                                mem_derivs_f32x2(xs2, cm1, dxs1, dxl1);                                                  // :84-85
                                const float2 xs_f = euler_clamp_f32x2(xs2, dxs1, dt2, Kc<float>::EPSILON, (float)hi_s);  // :125
                                const float2 xl_f = euler_clamp_f32x2(xl2, dxl1, dt2, 1.0f, (float)a.xl_max);
                                const float2 xs_h = euler_clamp_f32x2(xs2, dxs1, h2, Kc<float>::EPSILON, (float)hi_s);   // :128
                                const float2 xl_h = euler_clamp_f32x2(xl2, dxl1, h2, 1.0f, (float)a.xl_max);
                                const float2 mn = clause_rhs_f32x2(v2, d2, qf, xs_h, xl_h);                              // :129
                                mem_derivs_f32x2(xs_h, mul2(bc2(0.5f), mn), dxs2, dxl2);
                                const float2 xs_n = euler_clamp_f32x2(xs_h, dxs2, h2, Kc<float>::EPSILON, (float)hi_s);  // :130
                                const float2 xl_n = euler_clamp_f32x2(xl_h, dxl2, h2, 1.0f, (float)a.xl_max);
                                const float es[2] = {fabsf(__fsub_rn(xs_f.x, xs_n.x)), fabsf(__fsub_rn(xs_f.y, xs_n.y))};   // :104-107
                                const float el[2] = {fabsf(__fsub_rn(xl_f.x, xl_n.x)), fabsf(__fsub_rn(xl_f.y, xl_n.y))};
                                const float ns[2] = {xs_n.x, xs_n.y}, nl[2] = {xl_n.x, xl_n.y};
#pragma unroll
                                for (int w = 0; w < W; ++w) {
                                    if (!frozen[w]) {
                                        if (es[w] == es[w]) { const U b = (U)EB::enc((T)es[w]); e_loc[w] = b > e_loc[w] ? b : e_loc[w]; }
                                        if (el[w] == el[w]) { const U b = (U)EB::enc((T)el[w]); e_loc[w] = b > e_loc[w] ? b : e_loc[w]; }
                                        xs[w] = (T)ns[w];
                                        xl[w] = (T)nl[w];
                                    }
                                }
                                __stcg(at16(my_mem, it.x), IO::pack_mem(xs, xl));
                            }
#pragma unroll
                            for (int j = 0; j < 3; ++j) { d[j][0] = d2[j].x; d[j][1] = d2[j].y; }
                        }
                        if (!loopc) {   // (a loop clause has written its rows itself)
                            IO::store_dv(r0, d[0]);
                            if (!no1) IO::store_dv(r1, d[1]);
                            if (!no2) IO::store_dv(r2, d[2]);
                        }
                    }
                    {   // refill stage k with item i + D, wrapping into the other pass (of the next step after pass B)
                        int nx = i + D;
                        bool nb = pass != 0;
                        if (nx >= n_items) { nx -= n_items; nb = !nb; }
                        fetch(k, s_items[nx], nb);
                    }
                    if ((int)it.y < 0) __syncthreads();         // last item of a level: block-uniform
                }
            }
            if (pass == 0) {
                // -------------------- flag (:120-122) + variable pass A ---------------------------
                if constexpr (!STRICT && W == 2 && sizeof(T) == 4) { unsat[0] = unsat[0] || !(mx[0] < 0.5f); unsat[1] = unsat[1] || !(mx[1] < 0.5f); }
                unsigned any_unsat = 0;
#pragma unroll
                for (int w = 0; w < W; ++w) any_unsat |= (__syncthreads_or((int)unsat[w]) ? 1u : 0u) << w;
#pragma unroll
                for (int w = 0; w < W; ++w) {
                    if (!frozen[w] && !((any_unsat >> w) & 1u)) {   // all satisfied: state untouched, the loop of `simulate` ends
                        solved_at[w] = a.step0 + s;
                        if (tid == 0) a.solved[tile * W + w] = solved_at[w];
                        frozen[w] = true;
                    }
                }
                for (int i = tid; i < a.N; i += NT) {
                    T v[W], dv[W], vf[W];
                    IO::unpack(rows[i], v, dv);
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                        vf[w] = frozen[w] ? v[w] : euler_clamp(v[w], dv[w], dtw[w], T(-1), T(1));   // :125
                        v[w] = frozen[w] ? v[w] : euler_clamp(v[w], dv[w], hw[w], T(-1), T(1));     // :128
                        dv[w] = T(0);
                    }
                    __stcg(reinterpret_cast<uint2*>(vfull) + i, pack_cm(vf));
                    rows[i] = IO::pack(v, dv);
                }
                __syncthreads();
            } else {
                // -------------------- variable pass B (:130) + error norm + dt (:132-135) ---------
                // v_full comes back from L2: four rows' loads are issued before the first is used (ncu: a single dependent
                // load per iteration left 7 % of the kernel's stall samples on the subtraction below)
                for (int i0 = tid; i0 < a.N; i0 += 4 * NT) {
                    uint2 vfu[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * NT;
                        vfu[u] = i < a.N ? __ldcg(reinterpret_cast<const uint2*>(vfull) + i) : make_uint2(0u, 0u);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * NT;
                        if (i < a.N) {
                            T v[W], dv[W], vf[W];
                            IO::unpack(rows[i], v, dv);
                            unpack_cm(vfu[u], vf);
#pragma unroll
                            for (int w = 0; w < W; ++w) {
                                if (!frozen[w]) {
                                    v[w] = euler_clamp(v[w], dv[w], hw[w], T(-1), T(1));
                                    const T e = fabs(vf[w] - v[w]);                                 // :102-103
                                    if (e == e) { const U b = EB::enc(e); e_loc[w] = b > e_loc[w] ? b : e_loc[w]; }
                                }
                                dv[w] = T(0);
                            }
                            rows[i] = IO::pack(v, dv);
                        }
                    }
                }
#pragma unroll
                for (int w = 0; w < W; ++w) {
                    const U m = warp_max_bits<U>(e_loc[w]);
                    if ((tid & 31u) == 0u && m != EB::NONE) atomicMax(&s_err[w], m);
                }
                __syncthreads();
#pragma unroll
                for (int w = 0; w < W; ++w) {
                    if (!frozen[w]) {
                        const T e = EB::dec(s_err[w]);
                        dtw[w] = rmax(rmin(dtw[w] * sqrt(aa.tol / e), T(1e3)), T(0.0078125));       // :133-135
                    }
                }
                __syncthreads();
                if (tid < W) s_err[tid] = EB::NONE;     // next written in the NEXT step's variable pass B, many barriers away
            }
        }
    }
    cp_async_wait<0>();
    for (int i = tid; i < a.N; i += NT) {
        T v[W], dv[W];
        IO::unpack(rows[i], v, dv);
#pragma unroll
        for (int w = 0; w < W; ++w) vt[(int64_t)i * W + w] = v[w];
    }
    if (tid == 0) {
#pragma unroll
        for (int w = 0; w < W; ++w) if (valid[w]) aa.dt[tile * W + w] = dtw[w];
    }
}

}  // namespace odesat
