// tile_ragged.cuh — RAGGED clause lengths in the tile engine (included by tile_engine.cuh).
//
// The reference's clauses have any length: `solve -r` preprocessing produces resolvents of every size
// (cnf.rs:397-416), DIMACS files mix binary, ternary and long clauses, and SURVEY quirk Q9 makes unit and empty clauses
// part of the specification (second_min = +inf, C_m = +inf).  Round 1's tile kernels hold exactly three literals in a
// packed 8-byte word.  This kernel is k_tile_fixed (same rows in shared memory, same level schedule, same per-thread
// cp.async ring, same flags / freezing) with a second kind of slot: a LOOP clause, whose entry holds {offset, length}
// into an array of literal words (tile_schedule.hpp, TileSchedule::aux) and whose thread walks the literals twice —
// once for min / second-min (system.rs:45-58), once for the contributions (system.rs:62-81) — executing the reference's
// statements literally, so every length (0 and 1 included) yields the reference's bits without any domain argument.
// Clauses of one to three literals (distinct variables) keep the packed word, the bank-conflict packing and clause_math;
// a literal position that does not exist enters with the value +inf (system.rs:46-47).  Under the EXACT schedule clauses
// of 4..32 distinct literals are GROUP clauses, one lane per literal (clause_group).  The schedule compiler puts a
// level's group clauses first and its loop clauses last, so the kinds mostly sit in different warps, and a level
// still never contains two clauses with a common variable (the levels are built from the clauses' real literal lists).
#pragma once

namespace odesat {

// One LOOP clause for the W replicas of the tile: system.rs:43-88 statement by statement + the update (:94-95).
template <typename T, int W>
__device__ __forceinline__ void clause_loop(unsigned char* smem_raw, const uint32_t* __restrict__ lits, unsigned len, T (&xs)[W], T (&xl)[W],
                                            const bool (&frozen)[W], bool (&unsat)[W], T dt, T zeta, T xl_max) {
    using Row = typename TileTraits<T>::Row;
    using IO = RowIO<T, W>;
    const T hi_s = T(1) - Kc<T>::EPSILON;
    T mn[W], sm[W];
#pragma unroll
    for (int w = 0; w < W; ++w) { mn[w] = inf_v<T>(); sm[w] = inf_v<T>(); }
    // the literal words are read as 16-byte vectors, eight literals per round trip (a clause's list starts on a
    // 16-byte boundary); the second sweep finds them in L1
    const uint4* lits4 = reinterpret_cast<const uint4*>(lits);
    for (unsigned j0 = 0; j0 < len; j0 += 8) {
        const uint4 wa = __ldg(lits4 + (j0 >> 2));
        const uint4 wb = j0 + 4 < len ? __ldg(lits4 + (j0 >> 2) + 1) : make_uint4(0u, 0u, 0u, 0u);
        const uint32_t lws[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (j0 + u < len) {
                const uint32_t lw = lws[u];
                const Row* r = reinterpret_cast<const Row*>(smem_raw + (lw & 0x3FFF0u));
                const T q = (lw >> 31) ? T(-1) : T(1);
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
    T cm[W], wgt[W], rg[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
        cm[w] = T(0.5) * mn[w];                                                             // :60
        wgt[w] = xl[w] * xs[w];
        rg[w] = (T(1) + zeta * xl[w]) * (T(1) - xs[w]);
    }
    for (unsigned j0 = 0; j0 < len; j0 += 8) {
        const uint4 wa = __ldg(lits4 + (j0 >> 2));
        const uint4 wb = j0 + 4 < len ? __ldg(lits4 + (j0 >> 2) + 1) : make_uint4(0u, 0u, 0u, 0u);
        const uint32_t lws[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (j0 + u < len) {   // in literal order: a variable repeated inside the clause sees its own earlier addend
                const uint32_t lw = lws[u];
                Row* r = reinterpret_cast<Row*>(smem_raw + (lw & 0x3FFF0u));
                const T q = (lw >> 31) ? T(-1) : T(1);
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
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const T dxs = (Kc<T>::BETA * (xs[w] + Kc<T>::EPSILON)) * (cm[w] - Kc<T>::GAMMA);   // :84
        const T dxl = Kc<T>::ALPHA * (cm[w] - Kc<T>::DELTA);                               // :85
        unsat[w] = unsat[w] || !(cm[w] < Kc<T>::GAMMA);                                    // :88
        if (!frozen[w]) {
            xs[w] = euler_clamp(xs[w], dxs, dt, Kc<T>::EPSILON, hi_s);                     // :94
            xl[w] = euler_clamp(xl[w], dxl, dt, T(1), xl_max);                             // :95
        }
    }
}

// A loop clause of at most eight literals with distinct variables (TILE_ENTRY_LOOP8): the same statements, but all rows
// are taken into registers at once — one shared-memory round trip instead of two dependent ones per literal (the general
// form must re-read a row after every store because a repeated variable sees its own earlier addend).
template <typename T, int W>
__device__ __forceinline__ void clause_loop8(unsigned char* smem_raw, const uint32_t* __restrict__ lits, unsigned len, T (&xs)[W], T (&xl)[W],
                                             const bool (&frozen)[W], bool (&unsat)[W], T dt, T zeta, T xl_max) {
    using Row = typename TileTraits<T>::Row;
    using IO = RowIO<T, W>;
    const T hi_s = T(1) - Kc<T>::EPSILON;
    const uint4* lits4 = reinterpret_cast<const uint4*>(lits);
    const uint4 wa = __ldg(lits4);
    const uint4 wb = len > 4 ? __ldg(lits4 + 1) : make_uint4(0u, 0u, 0u, 0u);
    const uint32_t lws[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
    T v[8][W], d[8][W], q[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        q[u] = (lws[u] >> 31) ? T(-1) : T(1);
        if (u < (int)len) IO::unpack(*reinterpret_cast<const Row*>(smem_raw + (lws[u] & 0x3FFF0u)), v[u], d[u]);
    }
    T mn[W], sm[W];
#pragma unroll
    for (int w = 0; w < W; ++w) { mn[w] = inf_v<T>(); sm[w] = inf_v<T>(); }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        if (u < (int)len) {
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const T a = T(1) - q[u] * v[u][w];                                              // :49
                if (a < mn[w]) { sm[w] = mn[w]; mn[w] = a; } else if (a < sm[w]) { sm[w] = a; }   // :50-55
            }
        }
    }
    T cm[W], wgt[W], rg[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
        cm[w] = T(0.5) * mn[w];                                                             // :60
        wgt[w] = xl[w] * xs[w];
        rg[w] = (T(1) + zeta * xl[w]) * (T(1) - xs[w]);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        if (u < (int)len) {
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const T a = T(1) - q[u] * v[u][w];
                const T g = (T(0.5) * q[u]) * ((a != mn[w]) ? mn[w] : sm[w]);              // :64-70
                const T rr = (cm[w] == a) ? T(0.5) * (q[u] - v[u][w]) : T(0);              // :73-77
                d[u][w] = d[u][w] + (wgt[w] * g + rg[w] * rr);                             // :80
            }
            IO::store_dv(reinterpret_cast<Row*>(smem_raw + (lws[u] & 0x3FFF0u)), d[u]);
        }
    }
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const T dxs = (Kc<T>::BETA * (xs[w] + Kc<T>::EPSILON)) * (cm[w] - Kc<T>::GAMMA);   // :84
        const T dxl = Kc<T>::ALPHA * (cm[w] - Kc<T>::DELTA);                               // :85
        unsat[w] = unsat[w] || !(cm[w] < Kc<T>::GAMMA);                                    // :88
        if (!frozen[w]) {
            xs[w] = euler_clamp(xs[w], dxs, dt, Kc<T>::EPSILON, hi_s);                     // :94
            xl[w] = euler_clamp(xl[w], dxl, dt, T(1), xl_max);                             // :95
        }
    }
}

// GROUP clause: four to 32 literals with distinct variables, ONE LANE PER LITERAL (tile_schedule.hpp: P consecutive
// slots aligned to P inside the warp).  Every lane evaluates its literal (system.rs:49), the group's min / second-min
// (:50-58 — order statistics of the non-NaN values, duplicates counted, so the fold order does not matter) come from a
// butterfly of warp shuffles, the leader's {xs, xl} are broadcast, every lane adds its own contribution (:62-81) and the
// leader updates the memories (:84-95).  Must be called by all 32 lanes of the warp (`grp` false for bystanders).
// → true for the leader (its cell is to be written back).
// Used under the EXACT schedule, where every level is one item and waits for its slowest thread: a long clause then
// costs one literal's latency instead of a chain of them (measurements: tile_schedule.hpp, tile_use_groups).
template <typename T, int W>
__device__ __forceinline__ bool clause_group(unsigned char* smem_raw, bool grp, const uint2 e, T (&xs)[W], T (&xl)[W], const bool (&frozen)[W],
                                             bool (&unsat)[W], T dt, T zeta, T xl_max) {
    using Row = typename TileTraits<T>::Row;
    using IO = RowIO<T, W>;
    const T hi_s = T(1) - Kc<T>::EPSILON;
    const unsigned lane = threadIdx.x & 31u;
    const bool live = grp && !(e.y & TILE_ENTRY_VOID);
    const unsigned P = grp ? (1u << ((e.y >> 8) & 7u)) : 1u;
    const unsigned pos = grp ? (e.y & 31u) : 0u;
    Row* const r = reinterpret_cast<Row*>(smem_raw + (live ? (e.x & 0x3FFF0u) : 0u));
    const T q = (live && (e.x >> 31)) ? T(-1) : T(1);
    T v[W], d[W], a[W], mn[W], sm[W];
#pragma unroll
    for (int w = 0; w < W; ++w) { v[w] = T(0); d[w] = T(0); a[w] = inf_v<T>(); mn[w] = inf_v<T>(); sm[w] = inf_v<T>(); }
    if (live) {
        IO::unpack(*r, v, d);
#pragma unroll
        for (int w = 0; w < W; ++w) {
            a[w] = T(1) - q * v[w];                              // :49
            if (a[w] < mn[w]) mn[w] = a[w];                      // :50-52 on (inf, inf): a NaN never enters
        }
    }
#pragma unroll
    for (unsigned o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const T ymn = __shfl_xor_sync(0xFFFFFFFFu, mn[w], o), ysm = __shfl_xor_sync(0xFFFFFFFFu, sm[w], o);
            if (o < P) {                                         // partner inside this lane's group
                const T lo = rmin(mn[w], ymn), hi = rmax(mn[w], ymn);
                sm[w] = rmin(hi, rmin(sm[w], ysm));
                mn[w] = lo;
            }
        }
    }
    const int src = (int)((lane - pos) & 31u);                   // the group's leader
    T xsb[W], xlb[W];
#pragma unroll
    for (int w = 0; w < W; ++w) { xsb[w] = __shfl_sync(0xFFFFFFFFu, xs[w], src); xlb[w] = __shfl_sync(0xFFFFFFFFu, xl[w], src); }
    if (live) {
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const T cm = T(0.5) * mn[w];                                                    // :60
            const T wgt = xlb[w] * xsb[w];
            const T rg = (T(1) + zeta * xlb[w]) * (T(1) - xsb[w]);
            const T g = (T(0.5) * q) * ((a[w] != mn[w]) ? mn[w] : sm[w]);                  // :64-70
            const T rr = (cm == a[w]) ? T(0.5) * (q - v[w]) : T(0);                         // :73-77
            d[w] = d[w] + (wgt * g + rg * rr);                                              // :80
        }
        IO::store_dv(r, d);
    }
    if (!(live && pos == 0u)) return false;
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const T cm = T(0.5) * mn[w];
        const T dxs = (Kc<T>::BETA * (xs[w] + Kc<T>::EPSILON)) * (cm - Kc<T>::GAMMA);      // :84
        const T dxl = Kc<T>::ALPHA * (cm - Kc<T>::DELTA);                                  // :85
        unsat[w] = unsat[w] || !(cm < Kc<T>::GAMMA);                                       // :88
        if (!frozen[w]) {
            xs[w] = euler_clamp(xs[w], dxs, dt, Kc<T>::EPSILON, hi_s);                     // :94
            xl[w] = euler_clamp(xl[w], dxl, dt, T(1), xl_max);                             // :95
        }
    }
    return true;
}

// k_tile_fixed<T, NT, D, STRICT, ER = true, QUEUED = false> with loop clauses.  Shared memory as there:
// rows[N] (16 B) | ring_m[D][NT] (16 B) | ring_e[D][NT] (8 B) | items[n_items].
// GROUPS = false: the schedule holds no group clause (BALANCED, or no clause of 4..32 distinct literals) and the warp-wide
// group code is compiled out (measured: its presence alone cost the packed clauses 14 %, 0.66 → 0.76 ms/step on a 2/3-SAT mix).
template <typename T, int NT, int D, bool STRICT, bool GROUPS>
__global__ void __launch_bounds__(NT, 1) k_tile_ragged(const TileArgs<T> a, const uint32_t* __restrict__ aux) {
    constexpr int W = TileTraits<T>::W;
    using Row = typename TileTraits<T>::Row;
    using Mem = typename TileTraits<T>::Mem;
    using IO = RowIO<T, W>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Row* rows = reinterpret_cast<Row*>(smem_raw);
    Mem* ring_m = reinterpret_cast<Mem*>(smem_raw + (size_t)a.N * sizeof(Row));
    uint2* ring_e = reinterpret_cast<uint2*>(ring_m + D * NT);
    uint2* s_items = ring_e + D * NT;   // {slot base, count | last << 31}

    const int s_first = launch_first_step<STRICT>(a);   // block-uniform
    if (s_first >= a.nsteps) return;
    const unsigned tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int n_items = a.n_items;
    const uint2* my_entry = reinterpret_cast<const uint2*>(a.entry) + tid;      // + slot base
    Mem* my_mem = a.mem + tile * a.Mpad + tid;                                  // + slot base
    Mem* my_cell_m = ring_m + tid;                                              // + k·NT
    uint2* my_cell_e = ring_e + tid;
    T* vt = a.vt + tile * a.N * W;

    for (int i = tid; i < n_items; i += NT) {
        const uint32_t it = a.items[i];
        s_items[i] = make_uint2(it & 0xFFFFFu, ((it >> 20) & 0x7FFu) | (it & TILE_ITEM_LAST));
    }
    for (int i = tid; i < a.N; i += NT) {
        T v[W], dv[W];
#pragma unroll
        for (int w = 0; w < W; ++w) { v[w] = vt[(int64_t)i * W + w]; dv[w] = T(0); }
        rows[i] = IO::pack(v, dv);
    }
    bool valid[W], frozen[W];
    int32_t solved_at[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
        valid[w] = tile * W + w < a.R;
        solved_at[w] = valid[w] ? a.solved[tile * W + w] : 0;
        frozen[w] = !valid[w] || (a.freeze && solved_at[w] >= 0);
    }
    __syncthreads();
    auto fetch = [&](int k, const uint2 it) {
        if (tid < (it.y & 0x7FFFFFFFu)) {
            cp_async16(my_cell_m + k * NT, at16(my_mem, it.x));
            cp_async8(my_cell_e + k * NT, at8(my_entry, it.x));
        }
        cp_async_commit();
    };
#pragma unroll
    for (int k = 0; k < D; ++k) fetch(k, s_items[k]);

    for (int s = s_first; s < a.nsteps; ++s) {
        bool all_frozen = true;
#pragma unroll
        for (int w = 0; w < W; ++w) all_frozen = all_frozen && frozen[w];
        if (all_frozen) break;
        bool unsat[W];
        float mx[2] = {0.0f, 0.0f};   // f32x2 fast path: running max of the packed clauses' minima (→ unsat)
        T dtw[W];   // fast 3-literal path: a frozen replica is integrated with dt = 0 (clause_math)
#pragma unroll
        for (int w = 0; w < W; ++w) { unsat[w] = false; dtw[w] = (!STRICT && frozen[w]) ? T(0) : a.dt; }
        // ------------------------------ clause phase -----------------------------------
        for (int base = 0; base < n_items; base += D) {
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const int i = base + k;
                const uint2 it = s_items[i];
                cp_async_wait<D - 1>();                     // this thread's cells of item i have landed
                const bool mine = tid < (it.y & 0x7FFFFFFFu);
                uint2 e = make_uint2(0u, 0u);
                T xs[W], xl[W];
#pragma unroll
                for (int w = 0; w < W; ++w) { xs[w] = T(0); xl[w] = T(0); }
                if (mine) {
                    e = my_cell_e[k * NT];
                    IO::unpack_mem(my_cell_m[k * NT], xs, xl);
                }
                const bool grp = GROUPS && mine && (e.y & TILE_ENTRY_GROUP) != 0u;
                bool write_back = false;
                if constexpr (GROUPS) {
                    if (__any_sync(0xFFFFFFFFu, grp))   // warp-uniform: the whole warp helps with the shuffles
                        write_back = clause_group<T, W>(smem_raw, grp, e, xs, xl, frozen, unsat, a.dt, a.zeta, a.xl_max);
                }
                if (mine && !grp) {
                    write_back = true;
                    if (e.y & TILE_ENTRY_LOOP) {
                        // (bit 26 is a sign bit in a packed entry: TILE_ENTRY_LOOP8 only means something here)
                        if (e.y & TILE_ENTRY_LOOP8)   // four to eight literals, distinct variables (or none at all)
                            clause_loop8<T, W>(smem_raw, aux + e.x, e.y & 0xFFFFu, xs, xl, frozen, unsat, a.dt, a.zeta, a.xl_max);
                        else                          // more than eight literals, or a repeated variable
                            clause_loop<T, W>(smem_raw, aux + e.x, e.y & 0xFFFFu, xs, xl, frozen, unsat, a.dt, a.zeta, a.xl_max);
                    } else {
                        Row* const r0 = reinterpret_cast<Row*>(smem_raw + (e.x & 0x3FFF0u));
                        Row* const r1 = reinterpret_cast<Row*>(smem_raw + ((e.x >> 14) & 0x3FFF0u));
                        Row* const r2 = reinterpret_cast<Row*>(smem_raw + (e.y & 0x3FFF0u));
                        const T q[3] = {(e.y >> 24) & 1u ? T(-1) : T(1), (e.y >> 25) & 1u ? T(-1) : T(1), (e.y >> 26) & 1u ? T(-1) : T(1)};
                        T v[3][W], d[3][W];
                        IO::unpack(*r0, v[0], d[0]);
                        IO::unpack(*r1, v[1], d[1]);
                        IO::unpack(*r2, v[2], d[2]);
                        // a literal position that does not exist (one- and two-literal clauses): value +inf — where min and
                        // second-min start (system.rs:46-47) — through v = −inf with q = +1; its row is not written
                        const bool no2 = (e.y & TILE_ENTRY_NO2) != 0u, no1 = (e.y & TILE_ENTRY_NO1) != 0u;
#pragma unroll
                        for (int w = 0; w < W; ++w) {
                            if (no2) v[2][w] = -inf_v<T>();
                            if (no1) v[1][w] = -inf_v<T>();
                        }
                        if constexpr (!STRICT && W == 2 && sizeof(T) == 4) {   // packed f32x2 arithmetic, as in k_tile_fixed
                            const float2 v2[3] = {make_float2(v[0][0], v[0][1]), make_float2(v[1][0], v[1][1]), make_float2(v[2][0], v[2][1])};
                            float2 d2[3] = {make_float2(d[0][0], d[0][1]), make_float2(d[1][0], d[1][1]), make_float2(d[2][0], d[2][1])};
                            float2 xs2 = make_float2(xs[0], xs[1]), xl2 = make_float2(xl[0], xl[1]);
                            const float qf[3] = {(float)q[0], (float)q[1], (float)q[2]};
                            clause_math_f32x2(v2, d2, qf, xs2, xl2, mx, make_float2(dtw[0], dtw[1]), a.xl_max);
#pragma unroll
                            for (int j = 0; j < 3; ++j) { d[j][0] = d2[j].x; d[j][1] = d2[j].y; }
                            xs[0] = xs2.x; xs[1] = xs2.y; xl[0] = xl2.x; xl[1] = xl2.y;
                        } else {
#pragma unroll
                            for (int w = 0; w < W; ++w) {
                                const T vv[3] = {v[0][w], v[1][w], v[2][w]};
                                T dd[3] = {d[0][w], d[1][w], d[2][w]};
                                clause_math<T, STRICT>(vv, dd, q, xs[w], xl[w], frozen[w], unsat[w], dtw[w], a.zeta, a.xl_max);
                                d[0][w] = dd[0]; d[1][w] = dd[1]; d[2][w] = dd[2];
                            }
                        }
                        IO::store_dv(r0, d[0]);
                        if (!no1) IO::store_dv(r1, d[1]);
                        if (!no2) IO::store_dv(r2, d[2]);
                    }
                }
                if (write_back) __stcg(at16(my_mem, it.x), IO::pack_mem(xs, xl));
                {   // refill stage k with item i + D (next step's item i + D − n_items at the end)
                    int nx = i + D;
                    if (nx >= n_items) nx -= n_items;
                    fetch(k, s_items[nx]);
                }
                if ((int)it.y < 0) __syncthreads();         // last item of a level: block-uniform
            }
        }
        // ------------------------------ flags + variable phase ---------------------------
        if constexpr (!STRICT && W == 2 && sizeof(T) == 4) { unsat[0] = unsat[0] || !(mx[0] < 0.5f); unsat[1] = unsat[1] || !(mx[1] < 0.5f); }
        unsigned any_unsat = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) any_unsat |= (__syncthreads_or((int)unsat[w]) ? 1u : 0u) << w;
        for (int i = tid; i < a.N; i += NT) {
            T v[W], dv[W];
            IO::unpack(rows[i], v, dv);
#pragma unroll
            for (int w = 0; w < W; ++w) {
                if (!frozen[w]) v[w] = euler_clamp(v[w], dv[w], a.dt, T(-1), T(1));   // :96
                dv[w] = T(0);
            }
            rows[i] = IO::pack(v, dv);
        }
#pragma unroll
        for (int w = 0; w < W; ++w) {
            if (valid[w] && !frozen[w] && !((any_unsat >> w) & 1u)) {
                // the pre-update state of this step was all-satisfied (system.rs:149-153)
                if (solved_at[w] < 0) {
                    solved_at[w] = a.step0 + s;
                    if (tid == 0) a.solved[tile * W + w] = solved_at[w];
                }
                if (a.freeze) frozen[w] = true;
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
    for (int i = tid; i < a.N; i += NT) {
        T v[W], dv[W];
        IO::unpack(rows[i], v, dv);
#pragma unroll
        for (int w = 0; w < W; ++w) vt[(int64_t)i * W + w] = v[w];
    }
}

}  // namespace odesat
