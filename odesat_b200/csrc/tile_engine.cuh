// tile_engine.cuh — the THROUGHPUT engine: variables of a replica tile live in shared memory.
//
// One CTA owns a tile of W replicas (W = 2 for f32, 1 for f64) for a whole chunk of Euler
// steps.  Its shared memory holds one 16-byte row {v[W], dv[W]} per variable (N = 10 000 →
// 160 KB of the 227 KB), so the three random reads of v per clause and the three accumulations
// into dv never leave the SM.  Per step the CTA streams the clauses once, in the level order
// compiled by tile_schedule.hpp: one thread = one clause for W replicas; it reads the packed
// clause (8 B, shared by every CTA → L2) and {xs[W], xl[W]} (16 B, coalesced, read-modify-write
// in place), gathers three rows with LDS.128, computes C_m, the gradient contributions and the
// memory updates in registers, and adds into dv with a plain shared-memory RMW — race-free
// because a level never contains two clauses with a common variable, and deterministic
// because levels are separated by block barriers.  After the last level the all-satisfied
// flag is a __syncthreads_or, and the variable pass applies v ← clamp(v + dt·dv), dv ← 0 in
// shared memory.  v touches HBM only at chunk boundaries; the per-step HBM stream is xs/xl.
//
// Layouts in HBM (tile-major so that a CTA's stream is contiguous):
//   vt  [tiles][N][W]            variables, tile layout
//   mem [tiles][Mpad][2W]        {xs[W], xl[W]} per clause SLOT (schedule order, padded)
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <string>

#include "common.cuh"
#include "formula.hpp"
#include "packed_f32x2.cuh"
#include "tile_schedule.hpp"

namespace odesat {

template <typename T> struct TileTraits;
template <> struct TileTraits<float> {
    static constexpr int W = 2;
    using Row = float4;   // v0 v1 dv0 dv1
    using Mem = float4;   // xs0 xs1 xl0 xl1
};
template <> struct TileTraits<double> {
    static constexpr int W = 1;
    using Row = double2;  // v dv
    using Mem = double2;  // xs xl
};

template <typename T> struct TileArgs {
    int64_t N = 0, Mpad = 0, R = 0;
    int n_items = 0;
    const uint32_t* items = nullptr;   // [n_items] packed (base, nvalid, last-of-level)
    const uint64_t* entry = nullptr;   // [Mpad] packed clauses
    T* vt = nullptr;
    typename TileTraits<T>::Mem* mem = nullptr;
    int32_t* solved = nullptr;
    T dt = T(0), zeta = T(0), xl_max = T(0);
    int32_t step0 = 0, nsteps = 0, freeze = 0;
    // `inter` early exit without a host round trip: the (all-reduced) key of the PREVIOUS chunk; when it names a
    // flagged replica this speculatively issued launch does nothing.
    const unsigned long long* stop_key = nullptr;
    // Device-side choice between the literal first step and the fast arithmetic: *oor != 0 ⇔ the imported state
    // holds a value outside the domain on which the two are bit-identical.  The first launch after an import is
    // issued as a PAIR: the STRICT kernel (one step; a no-op when *oor == 0) and the fast kernel (which skips the
    // step the STRICT kernel has taken when *oor != 0) — no device→host flag read in between.
    const unsigned* oor = nullptr;
};

constexpr unsigned long long KEY_NONE = 0x7FFFFFFFFFFFFFFFull;   // INT64_MAX: no replica has flagged

// Common prologue of the step kernels: → first step index this launch runs (nsteps = nothing to do).
template <bool STRICT, typename A> __device__ __forceinline__ int launch_first_step(const A& a) {
    if (a.stop_key != nullptr && *a.stop_key != KEY_NONE) return a.nsteps;
    if (a.oor != nullptr) {
        const unsigned o = *a.oor;
        if (STRICT) return o ? 0 : a.nsteps;
        return o ? 1 : 0;
    }
    return 0;
}

template <typename T, int W> struct RowIO;
template <> struct RowIO<float, 2> {
    __device__ static void unpack(const float4& r, float* v, float* dv) { v[0] = r.x; v[1] = r.y; dv[0] = r.z; dv[1] = r.w; }
    __device__ static float4 pack(const float* v, const float* dv) { return make_float4(v[0], v[1], dv[0], dv[1]); }
    // store only the dv half of a row (8-byte STS)
    __device__ static void store_dv(float4* row, const float* dv) { reinterpret_cast<float2*>(row)[1] = make_float2(dv[0], dv[1]); }
    __device__ static void unpack_mem(const float4& m, float* xs, float* xl) { xs[0] = m.x; xs[1] = m.y; xl[0] = m.z; xl[1] = m.w; }
    __device__ static float4 pack_mem(const float* xs, const float* xl) { return make_float4(xs[0], xs[1], xl[0], xl[1]); }
};
template <> struct RowIO<double, 1> {
    __device__ static void unpack(const double2& r, double* v, double* dv) { v[0] = r.x; dv[0] = r.y; }
    __device__ static double2 pack(const double* v, const double* dv) { return make_double2(v[0], dv[0]); }
    __device__ static void store_dv(double2* row, const double* dv) { reinterpret_cast<double*>(row)[1] = dv[0]; }
    __device__ static void unpack_mem(const double2& m, double* xs, double* xl) { xs[0] = m.x; xl[0] = m.y; }
    __device__ static double2 pack_mem(const double* xs, const double* xl) { return make_double2(xs[0], xl[0]); }
};

__device__ __forceinline__ float flip_sign(float x, unsigned neg) { return __uint_as_float(__float_as_uint(x) ^ (neg << 31)); }
__device__ __forceinline__ double flip_sign(double x, unsigned neg) {
    return __hiloint2double(__double2hiint(x) ^ (int)(neg << 31), __double2loint(x));
}
__device__ __forceinline__ float fma_exact(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_exact(double a, double b, double c) { return __fma_rn(a, b, c); }

// The arithmetic of one clause for one replica (system.rs:43-88), two code paths that return
// bit-identical results on the domain the integrator lives in:
//   STRICT  the reference's statements one by one (sequential min / second-min, literal rigidity
//           term).  Used for the first step after importing a state with some |v| > 1 or
//           non-finite, or with a non-finite zeta.
//   fast    valid when every v is finite in [-1, 1] (true after any clamp, system.rs:96) and
//           zeta is finite:
//             · 1 − q·v as one FMA (q = ±1 ⇒ the product is exact, so is the FMA)
//             · min / second-min as a 5-op min/max network (order statistics do not depend
//               on the evaluation order; no NaN can occur)
//             · 0.5·q·sel·(xl·xs) as ±((0.5·xl·xs)·sel): scaling by ±0.5 commutes with rounding
//             · the rigidity term dropped: it is ±0 and x + (±0) = x for every x the running
//               sum can hold
template <typename T, bool STRICT>
__device__ __forceinline__ void clause_math(const T (&v)[3], T (&d)[3], const T (&q)[3], T& xs, T& xl, bool frozen,
                                            bool& unsat, T dt, T zeta, T xl_max) {
    const T hi_s = T(1) - Kc<T>::EPSILON;
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
        // d + q·(h·sel): q = ±1 makes the product exact, so the FMA rounds exactly like d + (±x)
        const T h = T(0.5) * wgt;
#pragma unroll
        for (int j = 0; j < 3; ++j) d[j] = fma_exact(h * ((a[j] != mn) ? mn : sm), q[j], d[j]);
    }
    const T dxs = (Kc<T>::BETA * (xs + Kc<T>::EPSILON)) * (cm - Kc<T>::GAMMA);   // :84
    const T dxl = Kc<T>::ALPHA * (cm - Kc<T>::DELTA);                            // :85
    unsat = unsat || !(cm < Kc<T>::GAMMA);                                       // :88
    if (STRICT) {
        if (!frozen) {
            xs = euler_clamp(xs, dxs, dt, Kc<T>::EPSILON, hi_s);                 // :94
            xl = euler_clamp(xl, dxl, dt, T(1), xl_max);                         // :95
        }
    } else {
        // a frozen replica is integrated with dt = 0: its memories are finite and already inside
        // their clamp ranges, so y + 0·dy = y exactly and the clamps are the identity
        xs = euler_clamp(xs, dxs, dt, Kc<T>::EPSILON, hi_s);
        xl = euler_clamp(xl, dxl, dt, T(1), xl_max);
    }
}

// Asynchronous global→shared copies (LDGSTS) of the prefetch ring: no destination register and
// no scoreboard slot, so — unlike a ring of plain loads, whose LDGs all share one of the six
// per-warp scoreboards and therefore wait for each other — the copies really run D items ahead.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
// base + idx·16 / base + idx·8 as one IMAD.WIDE (FMA pipe) instead of a 4-instruction ALU sequence
template <typename P> __device__ __forceinline__ P* at16(P* base, unsigned idx) {
    unsigned long long r;
    asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(r) : "r"(idx), "l"((unsigned long long)base));
    return (P*)r;
}
template <typename P> __device__ __forceinline__ P* at8(P* base, unsigned idx) {
    unsigned long long r;
    asm("mad.wide.u32 %0, %1, 8, %2;" : "=l"(r) : "r"(idx), "l"((unsigned long long)base));
    return (P*)r;
}
// Read-only 8-byte load that stays where it is written: a volatile asm is not moved across the other
// volatile asms (cp.async wait / commit), so the packed clause of the NEXT item is really requested
// before this item's work — ptxas otherwise sinks the load to the end of the item (ncu: 11.6 % of the
// stall samples sat on its first use).
__device__ __forceinline__ uint2 ldg_nc_pinned(const uint2* p) {
    uint2 r;
    asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Work queue of the persistent kernels (k_tile_ws in tile_ws.cuh; k_tile_fixed when wk.counter != nullptr): work items
// are (sub-chunk of steps, tile) in sub-chunk-major order; see tile_ws.cuh.
struct TileWork {
    int* counter = nullptr;   // next work item (zeroed by the host before the launch)
    int* done = nullptr;      // [tiles] sub-chunks of the tile that are published (zeroed before the launch)
    int tiles = 0;
    int nsub = 1;             // sub-chunks per launch
    int ksub = 0;             // steps per sub-chunk
    int early = 0;            // 1: a warp releases its ring stage as soon as its cells are in registers (not at the end of the item)
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}


// One fused fixed Euler step per loop iteration, `nsteps` per launch; one CTA per replica tile.
//
// Work decomposition: the schedule is a sequence of ITEMS, each up to NT consecutive clause
// slots of one level; thread t takes slot base + t when t < nvalid.  Every thread walks the
// same item sequence, so the level barriers (after the last item of a level) are block-uniform.
//
// Latency hiding: a ring of D stages in shared memory, one 16-byte cell per thread and stage
// ({xs, xl} of the thread's slot, the HBM stream).  As soon as a thread has consumed its cell of
// item i it refills it with item i + D by cp.async (wrapping into the next Euler step), and
// waits with cp.async.wait_group D−1 before reading item i.  A cell and a clause slot are only
// ever touched by their own thread, so no barrier is involved.  NT·D·16 bytes are in flight per
// SM.  The packed clause (8 B, identical for every CTA, L2-resident) is loaded into registers one
// item ahead — a single outstanding LDG, so no scoreboard aliasing.
// The schedule guarantees n_items % D == 0 and n_items > D.
//
// Shared memory:  rows[N] (16 B) | ring_m[D][NT] (16 B) | items[n_items]
//
// ER (entry-in-ring): the packed clause word travels through the cp.async ring too (24-byte
// cells) instead of the 1-ahead register load.  Better when an item is short (narrow CTA): one
// item time does not always cover an L2 round trip.
// QUEUED = false: one CTA per tile, the whole launch (round 1; compiled without the work loop — wrapping the same code in
// a run-time loop cost the EXACT schedule 4.5 %: 0.645 → 0.675 ms/step).  QUEUED = true: the CTAs are persistent and take
// (sub-chunk, tile) work items from the queue like k_tile_ws, so that a shard with few tiles per SM (strong scaling)
// has no tail wave; the ring is re-primed per work item.
template <typename T, int NT, int D, bool STRICT, bool ER, bool QUEUED>
__global__ void __launch_bounds__(NT, 1) k_tile_fixed(const TileArgs<T> a, const TileWork wk) {
    constexpr int W = TileTraits<T>::W;
    using Row = typename TileTraits<T>::Row;
    using Mem = typename TileTraits<T>::Mem;
    using IO = RowIO<T, W>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Row* rows = reinterpret_cast<Row*>(smem_raw);
    Mem* ring_m = reinterpret_cast<Mem*>(smem_raw + (size_t)a.N * sizeof(Row));
    uint2* ring_e = reinterpret_cast<uint2*>(ring_m + D * NT);
    uint2* s_items = reinterpret_cast<uint2*>(ring_e + (ER ? D * NT : 0));   // {slot base, count | last << 31}

    const int s_first = launch_first_step<STRICT>(a);   // block-uniform
    if (s_first >= a.nsteps) return;
    const unsigned tid = threadIdx.x;
    const uint2* my_entry = reinterpret_cast<const uint2*>(a.entry) + tid;      // + slot base
    Mem* my_cell_m = ring_m + tid;                                              // + k·NT
    uint2* my_cell_e = ring_e + tid;
    const int n_items = a.n_items;
    __shared__ int s_work;

    for (int i = tid; i < n_items; i += NT) {
        const uint32_t it = a.items[i];
        s_items[i] = make_uint2(it & 0xFFFFFu, ((it >> 20) & 0x7FFu) | (it & TILE_ITEM_LAST));
    }
    constexpr bool queued = QUEUED;
    for (int round = 0; QUEUED || round == 0; ++round) {
    int sub = 0, sa = s_first, sb = a.nsteps;
    int64_t tile = blockIdx.x;
    if constexpr (QUEUED) {
        __syncthreads();
        if (tid == 0) s_work = atomicAdd(wk.counter, 1);
        __syncthreads();
        const int widx = s_work;
        if (widx >= wk.tiles * wk.nsub) break;
        sub = widx / wk.tiles;
        tile = widx - sub * wk.tiles;
        sa = s_first + sub * wk.ksub;
        sb = min(a.nsteps, sa + wk.ksub);
        if (sub > 0 && tid == 0) { while (ld_acquire_gpu(wk.done + tile) < sub) { } }
        __syncthreads();
    }
    if (sa < sb) {
    T* vt = a.vt + tile * a.N * W;
    Mem* my_mem = a.mem + tile * a.Mpad + tid;                                  // + slot base
    for (int i = tid; i < a.N; i += NT) {
        T v[W], dv[W];
#pragma unroll
        for (int w = 0; w < W; ++w) { v[w] = QUEUED ? __ldcg(vt + (int64_t)i * W + w) : vt[(int64_t)i * W + w]; dv[w] = T(0); }
        rows[i] = IO::pack(v, dv);
    }
    bool valid[W], frozen[W];
    int32_t solved_at[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
        valid[w] = tile * W + w < a.R;
        solved_at[w] = valid[w] ? (QUEUED ? __ldcg(a.solved + tile * W + w) : a.solved[tile * W + w]) : 0;
        frozen[w] = !valid[w] || (a.freeze && solved_at[w] >= 0);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const uint2 it = s_items[k];
        if (tid < (it.y & 0x7FFFFFFFu)) {
            cp_async16(my_cell_m + k * NT, at16(my_mem, it.x));
            if (ER) cp_async8(my_cell_e + k * NT, at8(my_entry, it.x));
        }
        cp_async_commit();
    }
    uint2 it_next = s_items[0];
    uint2 e_next = make_uint2(0u, 0u);
    if (!ER && tid < (it_next.y & 0x7FFFFFFFu)) e_next = __ldg(at8(my_entry, it_next.x));

    for (int s = sa; s < sb; ++s) {
        bool all_frozen = true;
#pragma unroll
        for (int w = 0; w < W; ++w) all_frozen = all_frozen && frozen[w];
        if (all_frozen) break;
        bool unsat[W];
        float mx[2] = {0.0f, 0.0f};   // f32x2 fast path: running max of the clause minima (→ unsat after the clause phase)
        T dtw[W];   // per-replica step: 0 freezes a replica without a branch (fast path only)
#pragma unroll
        for (int w = 0; w < W; ++w) { unsat[w] = false; dtw[w] = (!STRICT && frozen[w]) ? T(0) : a.dt; }
        if constexpr (sizeof(T) == 4) {   // keep the per-replica dt in registers (ptxas otherwise rebuilds it in every item)
#pragma unroll
            for (int w = 0; w < W; ++w) asm volatile("" : "+f"(dtw[w]));
        }
        // ------------------------------ clause phase -----------------------------------
        for (int base = 0; base < n_items; base += D) {
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const int i = base + k;
                const uint2 it = it_next;
                uint2 e = e_next;
                {   // descriptor (and, without ER, packed clause) of the NEXT item, wrapping into the next step
                    const int i1 = (i + 1 == n_items) ? 0 : i + 1;
                    it_next = s_items[i1];
                    if (!ER && tid < (it_next.y & 0x7FFFFFFFu)) e_next = ldg_nc_pinned(at8(my_entry, it_next.x));
                }
                cp_async_wait<D - 1>();                     // this thread's cell of item i has landed
                const bool mine = tid < (it.y & 0x7FFFFFFFu);
                if (ER && mine) e = my_cell_e[k * NT];
                if (mine) {
                    const Mem mm = my_cell_m[k * NT];
                    // byte offsets of the three rows (pre-shifted in the packed clause), signs in the top byte
                    Row* const r0 = reinterpret_cast<Row*>(smem_raw + (e.x & 0x3FFF0u));
                    Row* const r1 = reinterpret_cast<Row*>(smem_raw + ((e.x >> 14) & 0x3FFF0u));
                    Row* const r2 = reinterpret_cast<Row*>(smem_raw + (e.y & 0x3FFF0u));
                    const T q[3] = {(e.y >> 24) & 1u ? T(-1) : T(1), (e.y >> 25) & 1u ? T(-1) : T(1), (e.y >> 26) & 1u ? T(-1) : T(1)};
                    T v[3][W], d[3][W], xs[W], xl[W];
                    IO::unpack(*r0, v[0], d[0]);
                    IO::unpack(*r1, v[1], d[1]);
                    IO::unpack(*r2, v[2], d[2]);
                    IO::unpack_mem(mm, xs, xl);
                    if constexpr (!STRICT && W == 2 && sizeof(T) == 4) {
                        const float2 v2[3] = {make_float2(v[0][0], v[0][1]), make_float2(v[1][0], v[1][1]), make_float2(v[2][0], v[2][1])};
                        float2 d2[3] = {make_float2(d[0][0], d[0][1]), make_float2(d[1][0], d[1][1]), make_float2(d[2][0], d[2][1])};
                        float2 xs2 = make_float2(xs[0], xs[1]), xl2 = make_float2(xl[0], xl[1]);
                        clause_math_f32x2(v2, d2, q, xs2, xl2, mx, make_float2(dtw[0], dtw[1]), a.xl_max);
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
                    // only the dv half changes; measured on B200, the 8-byte store (with its 2-way bank
                    // conflict across the two octets of a half-warp) beats rewriting the full row
                    IO::store_dv(r0, d[0]);
                    IO::store_dv(r1, d[1]);
                    IO::store_dv(r2, d[2]);
                    __stcg(at16(my_mem, it.x), IO::pack_mem(xs, xl));
                }
                {   // refill cell k with item i + D (next step's item i + D − n_items at the end)
                    int nx = i + D;
                    if (nx >= n_items) nx -= n_items;
                    const uint2 itn = s_items[nx];
                    if (tid < (itn.y & 0x7FFFFFFFu)) {
                        cp_async16(my_cell_m + k * NT, at16(my_mem, itn.x));
                        if (ER) cp_async8(my_cell_e + k * NT, at8(my_entry, itn.x));
                    }
                    cp_async_commit();
                }
                if ((int)it.y < 0) __syncthreads();         // last item of a level: block-uniform
            }
        }
        // ------------------------------ flags + variable phase ---------------------------
        if constexpr (!STRICT && W == 2 && sizeof(T) == 4) { unsat[0] = !(mx[0] < 0.5f); unsat[1] = !(mx[1] < 0.5f); }
        unsigned any_unsat = 0;   // __syncthreads_or is a boolean OR: one call per replica of the tile
#pragma unroll
        for (int w = 0; w < W; ++w) any_unsat |= (__syncthreads_or((int)unsat[w]) ? 1u : 0u) << w;
        for (int i = tid; i < a.N; i += NT) {
            T v[W], dv[W];
            IO::unpack(rows[i], v, dv);
#pragma unroll
            for (int w = 0; w < W; ++w) {
                if (STRICT) { if (!frozen[w]) v[w] = euler_clamp(v[w], dv[w], a.dt, T(-1), T(1)); }   // :96
                else v[w] = euler_clamp(v[w], dv[w], dtw[w], T(-1), T(1));
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
                    if (tid == 0) { if (QUEUED) __stcg(a.solved + tile * W + w, solved_at[w]); else a.solved[tile * W + w] = solved_at[w]; }
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
        for (int w = 0; w < W; ++w) { if (QUEUED) __stcg(vt + (int64_t)i * W + w, v[w]); else vt[(int64_t)i * W + w] = v[w]; }
    }
    }   // sa < sb
    if constexpr (QUEUED) {
        if (wk.nsub > 1) {
            __threadfence();
            __syncthreads();
            if (tid == 0) st_release_gpu(wk.done + tile, sub + 1);
        }
    }
    }   // work items
}

// ---- TMA-fed variant (ODESAT_TILE_TMA=1) -----------------------------------------------------------------------
// Same algorithm as k_tile_fixed<…, ER = true>, but the ring is filled by the copy engine: after the barrier that
// ends an item, ONE thread arms the stage's mbarrier with the byte count and issues two bulk copies
// (cp.async.bulk, global → shared, complete_tx on the mbarrier) — the item's {xs, xl} cells (cnt·16 B) and its
// packed clause words (cnt·8 B, rounded up to 16) — instead of every thread issuing its own LDGSTS.  Consumers
// wait on the stage's mbarrier parity.  A stage is shared by the whole CTA, so every non-empty item ends with a
// block barrier (with one item per level, as the schedules are cut, that is the level barrier anyway), and every
// thread issues fence.proxy.async.global once per item so that the bulk read of a slot in the NEXT step sees this
// step's write-back (generic-proxy store → async-proxy read).
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
    asm volatile("{\n.reg .b64 t;\nmbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1;\n}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred P1;\nTMA_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra TMA_DONE;\nbra TMA_WAIT;\nTMA_DONE:\n}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <typename T, int NT, int D>
__global__ void __launch_bounds__(NT, 1) k_tile_fixed_tma(const TileArgs<T> a) {
    constexpr int W = TileTraits<T>::W;
    using Row = typename TileTraits<T>::Row;
    using Mem = typename TileTraits<T>::Mem;
    using IO = RowIO<T, W>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Row* rows = reinterpret_cast<Row*>(smem_raw);
    Mem* ring_m = reinterpret_cast<Mem*>(smem_raw + (size_t)a.N * sizeof(Row));
    uint2* ring_e = reinterpret_cast<uint2*>(ring_m + D * NT);
    uint2* s_items = reinterpret_cast<uint2*>(ring_e + D * NT);
    const int n_items = a.n_items;                             // a multiple of D (host-checked)
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(s_items + n_items + 2);

    const int s_first = launch_first_step<false>(a);   // block-uniform
    if (s_first >= a.nsteps) return;
    const unsigned tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    T* vt = a.vt + tile * a.N * W;
    Mem* tile_mem = a.mem + tile * a.Mpad;
    const uint2* entries = reinterpret_cast<const uint2*>(a.entry);

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
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < D; ++k) mbar_init(bars + k, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
    auto refill = [&](int k, uint2 it) {   // thread 0 only; stage k is free
        const unsigned cnt = it.y & 0x7FFFFFFFu;
        if (cnt == 0) return;
        const unsigned bm = cnt * (unsigned)sizeof(Mem), be = ((cnt + 1u) & ~1u) * 8u;
        mbar_expect_tx(bars + k, bm + be);
        bulk_g2s(ring_m + k * NT, tile_mem + it.x, bm, bars + k);
        bulk_g2s(ring_e + k * NT, entries + it.x, be, bars + k);
    };
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < D; ++k) refill(k, s_items[k]);
    }
    unsigned parity = 0;   // bit k: phase the consumers of stage k wait for next

    for (int s = s_first; s < a.nsteps; ++s) {
        bool all_frozen = true;
#pragma unroll
        for (int w = 0; w < W; ++w) all_frozen = all_frozen && frozen[w];
        if (all_frozen) break;
        bool unsat[W];
        float mx[2] = {0.0f, 0.0f};
        T dtw[W];
#pragma unroll
        for (int w = 0; w < W; ++w) { unsat[w] = false; dtw[w] = frozen[w] ? T(0) : a.dt; }
        for (int base = 0; base < n_items; base += D) {
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const int i = base + k;
                const uint2 it = s_items[i];
                const unsigned cnt = it.y & 0x7FFFFFFFu;     // block-uniform
                if (cnt != 0) {
                    // Cross-proxy ordering for the write-backs: a slot written here is read again by a bulk copy almost a
                    // whole step later; the fence is cumulative over this thread's earlier stores, so issuing it one item
                    // late (when the previous item's store has long completed) covers them without stalling at the barrier.
                    asm volatile("fence.proxy.async.global;" ::: "memory");
                    mbar_wait(bars + k, (parity >> k) & 1u);
                    parity ^= 1u << k;
                    if (tid < cnt) {
                        const Mem mm = ring_m[k * NT + tid];
                        const uint2 e = ring_e[k * NT + tid];
                        Row* const r0 = reinterpret_cast<Row*>(smem_raw + (e.x & 0x3FFF0u));
                        Row* const r1 = reinterpret_cast<Row*>(smem_raw + ((e.x >> 14) & 0x3FFF0u));
                        Row* const r2 = reinterpret_cast<Row*>(smem_raw + (e.y & 0x3FFF0u));
                        const T q[3] = {(e.y >> 24) & 1u ? T(-1) : T(1), (e.y >> 25) & 1u ? T(-1) : T(1), (e.y >> 26) & 1u ? T(-1) : T(1)};
                        T v[3][W], d[3][W], xs[W], xl[W];
                        IO::unpack(*r0, v[0], d[0]);
                        IO::unpack(*r1, v[1], d[1]);
                        IO::unpack(*r2, v[2], d[2]);
                        IO::unpack_mem(mm, xs, xl);
                        if constexpr (W == 2 && sizeof(T) == 4) {
                            const float2 v2[3] = {make_float2(v[0][0], v[0][1]), make_float2(v[1][0], v[1][1]), make_float2(v[2][0], v[2][1])};
                            float2 d2[3] = {make_float2(d[0][0], d[0][1]), make_float2(d[1][0], d[1][1]), make_float2(d[2][0], d[2][1])};
                            float2 xs2 = make_float2(xs[0], xs[1]), xl2 = make_float2(xl[0], xl[1]);
                            clause_math_f32x2(v2, d2, q, xs2, xl2, mx, make_float2(dtw[0], dtw[1]), a.xl_max);
#pragma unroll
                            for (int j = 0; j < 3; ++j) { d[j][0] = d2[j].x; d[j][1] = d2[j].y; }
                            xs[0] = xs2.x; xs[1] = xs2.y; xl[0] = xl2.x; xl[1] = xl2.y;
                        } else {
#pragma unroll
                            for (int w = 0; w < W; ++w) {
                                const T vv[3] = {v[0][w], v[1][w], v[2][w]};
                                T dd[3] = {d[0][w], d[1][w], d[2][w]};
                                clause_math<T, false>(vv, dd, q, xs[w], xl[w], frozen[w], unsat[w], dtw[w], a.zeta, a.xl_max);
                                d[0][w] = dd[0]; d[1][w] = dd[1]; d[2][w] = dd[2];
                            }
                        }
                        IO::store_dv(r0, d[0]);
                        IO::store_dv(r1, d[1]);
                        IO::store_dv(r2, d[2]);
                        __stcg(tile_mem + it.x + tid, IO::pack_mem(xs, xl));
                    }
                    __syncthreads();                         // the stage is consumed and the level's dv stores are complete
                }
                if (tid == 0) {                              // also after an empty item: its stage feeds item i + D
                    int nx = i + D;
                    if (nx >= n_items) nx -= n_items;
                    refill(k, s_items[nx]);
                }
            }
        }
        if constexpr (W == 2 && sizeof(T) == 4) { unsat[0] = !(mx[0] < 0.5f); unsat[1] = !(mx[1] < 0.5f); }
        unsigned any_unsat = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) any_unsat |= (__syncthreads_or((int)unsat[w]) ? 1u : 0u) << w;
        for (int i = tid; i < a.N; i += NT) {
            T v[W], dv[W];
            IO::unpack(rows[i], v, dv);
#pragma unroll
            for (int w = 0; w < W; ++w) { v[w] = euler_clamp(v[w], dv[w], dtw[w], T(-1), T(1)); dv[w] = T(0); }
            rows[i] = IO::pack(v, dv);
        }
#pragma unroll
        for (int w = 0; w < W; ++w) {
            if (valid[w] && !frozen[w] && !((any_unsat >> w) & 1u)) {
                if (solved_at[w] < 0) {
                    solved_at[w] = a.step0 + s;
                    if (tid == 0) a.solved[tile * W + w] = solved_at[w];
                }
                if (a.freeze) frozen[w] = true;
            }
        }
        __syncthreads();
    }
    // drain the copies that were requested for a step that does not run
#pragma unroll
    for (int k = 0; k < D; ++k)
        if ((s_items[k].y & 0x7FFFFFFFu) != 0u) mbar_wait(bars + k, (parity >> k) & 1u);
    for (int i = tid; i < a.N; i += NT) {
        T v[W], dv[W];
        IO::unpack(rows[i], v, dv);
#pragma unroll
        for (int w = 0; w < W; ++w) vt[(int64_t)i * W + w] = v[w];
    }
}

}  // namespace odesat
#include "tile_ws.cuh"
#include "tile_adaptive.cuh"
#include "tile_ragged.cuh"
#include "tile_adaptive_ws.cuh"
namespace odesat {

// ---- small-instance persistent kernel ---------------------------------------------
// When every level of the schedule fits in one warp (≤ 32 clauses, e.g. the reference's
// aim-100 fixtures: N = 100, M = 160) a replica tile is integrated by ONE WARP with the whole
// state resident in shared memory: variable rows, the {xs, xl} cells and the packed clause words
// are loaded once, the whole chunk of Euler steps runs without touching HBM, and the level
// barriers are __syncwarp().  Same arithmetic (clause_math) and same freeze semantics as
// k_tile_fixed.  Shared memory: rows[N] | cells[n_items][32] | entries[n_items][32] | items[n_items].
template <typename T, bool STRICT>
__global__ void __launch_bounds__(32) k_tile_small(const TileArgs<T> a) {
    constexpr int W = TileTraits<T>::W;
    using Row = typename TileTraits<T>::Row;
    using Mem = typename TileTraits<T>::Mem;
    using IO = RowIO<T, W>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n_items = a.n_items;
    Row* rows = reinterpret_cast<Row*>(smem_raw);
    Mem* cells = reinterpret_cast<Mem*>(smem_raw + (size_t)a.N * sizeof(Row));
    uint2* entries = reinterpret_cast<uint2*>(cells + (size_t)n_items * 32);
    uint2* s_items = entries + (size_t)n_items * 32;

    const int s_first = launch_first_step<STRICT>(a);
    if (s_first >= a.nsteps) return;
    const unsigned lane = threadIdx.x;
    const int64_t tile = blockIdx.x;
    T* vt = a.vt + tile * a.N * W;
    Mem* mem = a.mem + tile * a.Mpad;
    const uint2* entry = reinterpret_cast<const uint2*>(a.entry);

    for (int i = lane; i < n_items; i += 32) {
        const uint32_t it = a.items[i];
        s_items[i] = make_uint2(it & 0xFFFFFu, ((it >> 20) & 0x7FFu) | (it & TILE_ITEM_LAST));
    }
    for (int i = lane; i < a.N; i += 32) {
        T v[W], dv[W];
#pragma unroll
        for (int w = 0; w < W; ++w) { v[w] = vt[(int64_t)i * W + w]; dv[w] = T(0); }
        rows[i] = IO::pack(v, dv);
    }
    __syncwarp();
    for (int i = 0; i < n_items; ++i) {
        const uint2 it = s_items[i];
        if (lane < (it.y & 0x7FFFFFFFu)) {
            cells[i * 32 + lane] = mem[it.x + lane];
            entries[i * 32 + lane] = __ldg(entry + it.x + lane);
        }
    }
    bool valid[W], frozen[W];
    int32_t solved_at[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
        valid[w] = tile * W + w < a.R;
        solved_at[w] = valid[w] ? a.solved[tile * W + w] : 0;
        frozen[w] = !valid[w] || (a.freeze && solved_at[w] >= 0);
    }
    __syncwarp();

    for (int s = s_first; s < a.nsteps; ++s) {
        bool all_frozen = true;
#pragma unroll
        for (int w = 0; w < W; ++w) all_frozen = all_frozen && frozen[w];
        if (all_frozen) break;
        bool unsat[W];
        float mx[2] = {0.0f, 0.0f};
        T dtw[W];
#pragma unroll
        for (int w = 0; w < W; ++w) { unsat[w] = false; dtw[w] = (!STRICT && frozen[w]) ? T(0) : a.dt; }
        for (int i = 0; i < n_items; ++i) {
            const uint2 it = s_items[i];
            if (lane < (it.y & 0x7FFFFFFFu)) {
                const uint2 e = entries[i * 32 + lane];
                const Mem mm = cells[i * 32 + lane];
                Row* const r0 = reinterpret_cast<Row*>(smem_raw + (e.x & 0x3FFF0u));
                Row* const r1 = reinterpret_cast<Row*>(smem_raw + ((e.x >> 14) & 0x3FFF0u));
                Row* const r2 = reinterpret_cast<Row*>(smem_raw + (e.y & 0x3FFF0u));
                const T q[3] = {(e.y >> 24) & 1u ? T(-1) : T(1), (e.y >> 25) & 1u ? T(-1) : T(1), (e.y >> 26) & 1u ? T(-1) : T(1)};
                T v[3][W], d[3][W], xs[W], xl[W];
                IO::unpack(*r0, v[0], d[0]);
                IO::unpack(*r1, v[1], d[1]);
                IO::unpack(*r2, v[2], d[2]);
                IO::unpack_mem(mm, xs, xl);
                if constexpr (!STRICT && W == 2 && sizeof(T) == 4) {
                    const float2 v2[3] = {make_float2(v[0][0], v[0][1]), make_float2(v[1][0], v[1][1]), make_float2(v[2][0], v[2][1])};
                    float2 d2[3] = {make_float2(d[0][0], d[0][1]), make_float2(d[1][0], d[1][1]), make_float2(d[2][0], d[2][1])};
                    float2 xs2 = make_float2(xs[0], xs[1]), xl2 = make_float2(xl[0], xl[1]);
                    clause_math_f32x2(v2, d2, q, xs2, xl2, mx, make_float2(dtw[0], dtw[1]), a.xl_max);
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
                IO::store_dv(r1, d[1]);
                IO::store_dv(r2, d[2]);
                cells[i * 32 + lane] = IO::pack_mem(xs, xl);
            }
            if ((int)it.y < 0) __syncwarp();                // last item of a level
        }
        if constexpr (!STRICT && W == 2 && sizeof(T) == 4) { unsat[0] = !(mx[0] < 0.5f); unsat[1] = !(mx[1] < 0.5f); }
        unsigned any_unsat = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) any_unsat |= (__any_sync(0xFFFFFFFFu, unsat[w]) ? 1u : 0u) << w;
        for (int i = lane; i < a.N; i += 32) {
            T v[W], dv[W];
            IO::unpack(rows[i], v, dv);
#pragma unroll
            for (int w = 0; w < W; ++w) {
                if (STRICT) { if (!frozen[w]) v[w] = euler_clamp(v[w], dv[w], a.dt, T(-1), T(1)); }   // :96
                else v[w] = euler_clamp(v[w], dv[w], dtw[w], T(-1), T(1));
                dv[w] = T(0);
            }
            rows[i] = IO::pack(v, dv);
        }
#pragma unroll
        for (int w = 0; w < W; ++w) {
            if (valid[w] && !frozen[w] && !((any_unsat >> w) & 1u)) {
                if (solved_at[w] < 0) {
                    solved_at[w] = a.step0 + s;
                    if (lane == 0) a.solved[tile * W + w] = solved_at[w];
                }
                if (a.freeze) frozen[w] = true;
            }
        }
        __syncwarp();
    }
    for (int i = lane; i < a.N; i += 32) {
        T v[W], dv[W];
        IO::unpack(rows[i], v, dv);
#pragma unroll
        for (int w = 0; w < W; ++w) vt[(int64_t)i * W + w] = v[w];
    }
    for (int i = 0; i < n_items; ++i) {
        const uint2 it = s_items[i];
        if (lane < (it.y & 0x7FFFFFFFu)) mem[it.x + lane] = cells[i * 32 + lane];
    }
}

// canonical replica-major [row][Rp]  →  tile layouts
template <typename T>
__global__ void k_tile_import(const T* __restrict__ v, const T* __restrict__ xs, const T* __restrict__ xl, int64_t Rp, int64_t R,
                              int64_t N, int64_t Mpad, const int32_t* __restrict__ perm, T* __restrict__ vt,
                              typename TileTraits<T>::Mem* __restrict__ mem, int64_t tiles, unsigned* __restrict__ out_of_range) {
    constexpr int W = TileTraits<T>::W;
    using IO = RowIO<T, W>;
    const int64_t tile = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t row = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    if (tile >= tiles || row >= N + Mpad) return;
    if (row < N) {
        bool bad = false;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const int64_t r = tile * W + w;
            const T x = r < R ? v[row * Rp + r] : T(0);
            vt[(tile * N + row) * W + w] = x;
            bad = bad || !(fabs(x) <= T(1));   // |v| > 1 or NaN
        }
        if (bad) *out_of_range = 1u;
    } else {
        const int64_t slot = row - N;
        const int m = perm[slot];
        T a[W], b[W];
        bool bad = false;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const int64_t r = tile * W + w;
            const bool ok = m >= 0 && r < R;
            a[w] = ok ? xs[(int64_t)m * Rp + r] : T(0);
            b[w] = ok ? xl[(int64_t)m * Rp + r] : T(0);
            bad = bad || !mem_in_fast_domain(a[w]) || !mem_in_fast_domain(b[w]);
        }
        if (bad) *out_of_range = 1u;
        mem[tile * Mpad + slot] = IO::pack_mem(a, b);
    }
}

template <typename T>
__global__ void k_tile_export(T* __restrict__ v, T* __restrict__ xs, T* __restrict__ xl, int64_t Rp, int64_t R, int64_t N,
                              int64_t Mpad, const int32_t* __restrict__ perm, const T* __restrict__ vt,
                              const typename TileTraits<T>::Mem* __restrict__ mem, int64_t tiles) {
    constexpr int W = TileTraits<T>::W;
    using IO = RowIO<T, W>;
    const int64_t tile = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t row = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    if (tile >= tiles || row >= N + Mpad) return;
    if (row < N) {
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const int64_t r = tile * W + w;
            if (r < R) v[row * Rp + r] = vt[(tile * N + row) * W + w];
        }
    } else {
        const int64_t slot = row - N;
        const int m = perm[slot];
        if (m < 0) return;
        T a[W], b[W];
        IO::unpack_mem(mem[tile * Mpad + slot], a, b);
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const int64_t r = tile * W + w;
            if (r < R) { xs[(int64_t)m * Rp + r] = a[w]; xl[(int64_t)m * Rp + r] = b[w]; }
        }
    }
}

// ---- tile-layout shortcuts of the e2e path (no canonical round trip) ----------------------------------
// xs = xs0 (system.rs:362-372), xl = 1 (main.rs:287) for every replica, written straight into the slot order.
template <typename T>
__global__ void k_tile_init_mem(const int8_t* __restrict__ xs0, const int32_t* __restrict__ perm,
                                typename TileTraits<T>::Mem* __restrict__ mem, int64_t Mpad, int64_t tiles) {
    constexpr int W = TileTraits<T>::W;
    using IO = RowIO<T, W>;
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t tile = blockIdx.y;
    if (slot >= Mpad || tile >= tiles) return;
    const int m = perm[slot];
    T a[W], b[W];
#pragma unroll
    for (int w = 0; w < W; ++w) { a[w] = m >= 0 ? (T)xs0[m] : T(0); b[w] = m >= 0 ? T(1) : T(0); }
    mem[tile * Mpad + slot] = IO::pack_mem(a, b);
}
// canonical v[row][Rp] → vt[tile][row][W] through a shared-memory transpose (both sides coalesced)
template <typename T>
__global__ void k_tile_import_v(const T* __restrict__ v, int64_t Rp, int64_t R, int64_t N, T* __restrict__ vt, int64_t tiles,
                                unsigned* __restrict__ out_of_range) {
    constexpr int W = TileTraits<T>::W;
    __shared__ T t[32][33];
    const int64_t x0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    bool bad = false;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int64_t row = x0 + k, r = r0 + threadIdx.x;
        const T x = (row < N && r < R) ? v[row * Rp + r] : T(0);
        t[k][threadIdx.x] = x;
        bad = bad || !(fabs(x) <= T(1));
    }
    if (bad) *out_of_range = 1u;
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int64_t r = r0 + k, row = x0 + threadIdx.x;
        if (row < N && r < tiles * W) vt[((r / W) * N + row) * W + (r % W)] = t[threadIdx.x][k];
    }
}
// cnf.rs:246-264 for the replicas of one tile with the thresholded v in shared memory (one byte per variable,
// bit w = v_w > 0); clauses come from the packed slot table.
template <typename T>
__global__ void __launch_bounds__(512) k_tile_verify(const T* __restrict__ vt, const uint64_t* __restrict__ entry,
                                                     const int32_t* __restrict__ perm, int64_t N, int64_t Mpad, int64_t R,
                                                     uint32_t* __restrict__ bad) {
    constexpr int W = TileTraits<T>::W;
    extern __shared__ __align__(16) unsigned char vbits[];
    const int64_t tile = blockIdx.x;
    const T* my = vt + tile * N * W;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        unsigned b = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) b |= (my[(int64_t)i * W + w] > T(0) ? 1u : 0u) << w;   // system.rs:238
        vbits[i] = (unsigned char)b;
    }
    __syncthreads();
    unsigned falsified = 0;   // bit w: some clause of replica w is falsified
    const uint2* e2 = reinterpret_cast<const uint2*>(entry);
    for (int64_t slot = threadIdx.x; slot < Mpad; slot += blockDim.x) {
        if (perm[slot] < 0) continue;
        const uint2 e = e2[slot];
        const unsigned i0 = (e.x & 0x3FFF0u) >> 4, i1 = ((e.x >> 14) & 0x3FFF0u) >> 4, i2 = (e.y & 0x3FFF0u) >> 4;
        const unsigned n0 = (e.y >> 24) & 1u ? 0xFFu : 0u, n1 = (e.y >> 25) & 1u ? 0xFFu : 0u, n2 = (e.y >> 26) & 1u ? 0xFFu : 0u;
        const unsigned sat = (vbits[i0] ^ n0) | (vbits[i1] ^ n1) | (vbits[i2] ^ n2);
        falsified |= ~sat;
    }
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const int any = __syncthreads_or((int)((falsified >> w) & 1u));
        if (threadIdx.x == 0 && tile * W + w < R) bad[tile * W + w] = any ? 1u : 0u;
    }
}
template <typename T>
__global__ void k_tile_assignment(const T* __restrict__ vt, int64_t N, int64_t rep, uint8_t* __restrict__ out) {
    constexpr int W = TileTraits<T>::W;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) out[i] = vt[((rep / W) * N + i) * W + (rep % W)] > T(0) ? 1 : 0;
}

// What BatchImpl needs from a shared-memory-resident engine (k_tile_fixed / k_tile_small here,
// the thread-block-cluster variant in tile_cluster.cuh).
template <typename T> struct TileBase {
    virtual ~TileBase() = default;
    virtual void reset_control() = 0;
    virtual int64_t import_state(const T* v, const T* xs, const T* xl, int64_t Rp) = 0;   // canonical [row][Rp] → tile layout
    virtual int64_t export_state(T* v, T* xs, T* xl, int64_t Rp) = 0;
    // enqueues n fixed steps (no host synchronisation); stop_key: see TileArgs; → launches
    virtual int64_t run_fixed(T dt, T zeta, int64_t n, int freeze, int32_t* solved, int64_t step0,
                              const unsigned long long* stop_key = nullptr) = 0;
    // adaptive steps (system.rs:111-139) with one dt per replica (dt_arr[R], device, in/out); every flagged replica is
    // frozen (its state untouched, :122).  Engines without it: has_adaptive() == false.
    virtual bool has_adaptive() const { return false; }
    virtual int64_t run_adaptive(T /*tol*/, T /*zeta*/, int64_t /*n*/, int32_t* /*solved*/, int64_t /*step0*/, T* /*dt_arr*/) {
        throw Error(ODESAT_EUNSUPPORTED, "this engine integrates fixed steps only; use the gather engine for adaptive steps");
    }
    // device-to-device copy of the whole tile-layout state, and back (lock-step `inter`: a chunk that overshot the
    // winning step is replayed from the chunk's start)
    virtual void snapshot() = 0;
    virtual void restore() = 0;
    // hint: another batch's kernels on the same device are enqueued right behind this one's (they fill its tail wave)
    virtual void set_followed(bool) {}
    // Optional shortcuts that work on the tile layout directly (no canonical round trip); 0 = not offered.
    virtual bool has_direct() const { return false; }
    virtual int64_t init_mem(const int8_t* /*xs0*/) { return 0; }                                   // xs = xs0, xl = 1 for every replica
    virtual int64_t import_v(const T* /*v*/, int64_t /*Rp*/) { return 0; }                          // canonical v only
    virtual int64_t verify_direct(uint32_t* /*bad*/) { return 0; }                                  // bad[r] = some clause falsified
    virtual int64_t assignment_direct(int64_t /*r*/, uint8_t* /*out_dev*/) { return 0; }            // out[i] = v_i > 0
};

template <typename T> struct TileEngine final : TileBase<T> {
    static constexpr int W = TileTraits<T>::W;
    using Mem = typename TileTraits<T>::Mem;
    static constexpr size_t kMaxSmem = 232448 - 1024;   // 227 KB opt-in limit minus static slack

    const odesat_formula& f;
    int64_t R, tiles;
    std::shared_ptr<TileSchedule> sched;
    cudaStream_t stream;
    DevBuf<T> vt, vt_snap;
    DevBuf<Mem> mem, mem_snap;
    DevBuf<unsigned> oor;        // device flag set by the import kernels: the state needs the literal first step
    bool need_rterm = true;      // the next launch is the first one after an import: issue the STRICT / fast pair
    bool oor_valid = false;      // *oor was written by an import since the last reset (else: assume the worst)
    int64_t* ledger_ = nullptr;
    // ring fed by cp.async.bulk (k_tile_fixed_tma).  Measured on B200 at the headline size, ms/step TMA vs per-thread
    // cp.async: BALANCED 768 threads f32 0.553 vs 0.589, f64 0.600 vs 0.629 (at 640) — the defaults below; f32 640
    // threads 0.592 vs 0.596; EXACT 512 threads 0.681 vs 0.636 — the elected-thread issue and the per-item barrier
    // only pay where an item is a full 768-clause level.  ODESAT_TILE_TMA=0/1 overrides.
    int tma_env = [] { const char* e = std::getenv("ODESAT_TILE_TMA"); return e ? (e[0] == '1' ? 1 : 0) : -1; }();
    bool use_tma = false;
    // warp-specialised persistent kernel (tile_ws.cuh): producer warp + full/empty mbarriers, barriers only between
    // levels, work queue of (sub-chunk, tile).  ODESAT_TILE_WS=0/1 overrides.
    int ws_env = [] { const char* e = std::getenv("ODESAT_TILE_WS"); return e ? (e[0] == '1' ? 1 : 0) : -1; }();
    // items per BALANCED level: 1 = one CTA width of clauses per level (a barrier after every item), 0 = wide levels,
    // as few colours as the maximum variable degree allows (tile_schedule.hpp).  ODESAT_TILE_IPL overrides.
    int ipl = tile_items_per_level();
    int ksub_env = [] { const char* e = std::getenv("ODESAT_TILE_KSUB"); return e ? std::atoi(e) : 0; }();
    bool use_ws = false;
    // the per-thread-ring kernel (EXACT, and every fallback) can take work from the same queue when that removes a tail
    // wave.  Measured on B200 (EXACT, 256 tiles on 148 SMs, 20 steps): f64 0.1033 → 0.0982 ms/step, f32 0.0951 → 0.0982
    // (the queued instantiation's code is ≈ 4.5 % slower per step, which eats the f32 gain) — default: f64 only.
    // ODESAT_TILE_QUEUE=0/1 overrides.
    bool queue_fixed = [] { const char* e = std::getenv("ODESAT_TILE_QUEUE"); return e ? e[0] != '0' : sizeof(T) == 8; }();
    int num_sms = 148;
    DevBuf<int> work;     // [1 + tiles] work-queue counter and per-tile published sub-chunks (k_tile_ws)
    bool small = false;   // one warp per tile, state resident in shared memory (k_tile_small)
    bool ragged = false;  // some clause is not three distinct variables: k_tile_ragged (tile_ragged.cuh), fixed steps only
    int nt = 512;
    int chunk = 64;   // Euler steps per launch
    int depth = 6;    // prefetch ring depth (tunable for NT = 512 only)

    static size_t smem_small(int64_t N, int n_items) { return (size_t)N * 16 + (size_t)n_items * (32 * 24 + 8); }
    static bool entry_in_ring(int nt) { return nt < 1024; }
    static size_t smem_bytes(int64_t N, int n_items, int nt, int depth) {
        return (size_t)N * 16 + (size_t)nt * depth * (entry_in_ring(nt) ? 24 : 16) + (size_t)(n_items + 2) * 8;
    }
    // deepest ring (≤ 6) that fits beside the variable rows
    static int pick_depth(int64_t N, int nt, int n_items_guess) {
        for (int d = 6; d >= 2; --d)
            if (smem_bytes(N, n_items_guess, nt, d) <= kMaxSmem) return d;
        return 0;
    }

    static bool supports(const odesat_formula& f, int64_t R, std::string* why) {
        auto no = [&](const char* m) { if (why) *why = m; return false; };
        if (R < 1) return no("empty batch");
        // uniform 3-literal clauses with distinct variables: the packed kernels; anything else (ragged lengths, unit and
        // empty clauses, a variable repeated inside a clause): k_tile_ragged, fixed steps
        if (f.M < 1) return no("no clauses");
        if (f.N > 16383) return no("more than 16383 variables");
        // a 512-thread CTA with a ring of 2 and an item table sized for this formula must fit beside the rows
        // (a narrower CTA would fit a few more variables but cannot hide the shared-memory latency: the
        // one-replica kernel of tile_cluster.cuh takes over from there)
        if (pick_depth(f.N, 512, (int)(f.M / 512 + 4 * f.max_degree + 64)) < 2) return no("variables do not fit in 227 KB of shared memory");
        return true;
    }
    static bool preferred(const odesat_formula&, int64_t R) { return R >= 8; }

    TileEngine(const odesat_formula& f_, int64_t R_, int kind, cudaStream_t st, int64_t* ledger) : f(f_), R(R_), stream(st), ledger_(ledger) {
        tiles = (R + W - 1) / W;
        ragged = f.K != 3 || !f.distinct_vars;
        if (ragged) { tma_env = 0; ws_env = 0; queue_fixed = false; }
        // levels: BALANCED colour classes do not depend on the CTA width; EXACT levels are list-scheduled
        // with the CTA width as the cap (one item per level), so they are built per candidate width
        auto levels_for = [&](int cap) {
            const int wide = (kind == ODESAT_SCHED_BALANCED && cap >= 512 && ipl != 1) ? 1 + ipl : 0;
            const int key = kind + 2 * cap + 4096 * wide;
            auto it = f.tile_levels.find(key);
            if (it == f.tile_levels.end()) {
                // BALANCED: colour classes of one item (a CTA width of clauses), capacity rounded to half an item —
                // or, wide, of several whole items
                it = f.tile_levels.emplace(key, kind == ODESAT_SCHED_EXACT ? build_tile_levels(f, kind, cap)
                                                                           : build_balanced_levels(f, cap, ipl)).first;
            }
            return it->second;
        };
        static const int cand[] = {128, 512, 640, 704, 768, 1024};
        if (const char* e = std::getenv("ODESAT_TILE_CHUNK")) { const int v = std::atoi(e); if (v > 0) chunk = v; }
        std::shared_ptr<TileLevels> lv;
        {   // small-instance mode: every level fits in a warp and the whole tile state fits in shared memory
            const char* e = std::getenv("ODESAT_TILE_SMALL");
            if (f.M <= 8192 && !ragged && !(e && e[0] == '0')) {
                auto l32 = levels_for(32);
                size_t maxlev = 0, items32 = 0;
                for (const auto& b : l32->bucket) { maxlev = std::max(maxlev, b.size()); items32 += b.empty() ? 0 : 1; }
                small = maxlev <= 32 && smem_small(f.N, (int)items32 + 2) <= kMaxSmem;
                if (small) lv = l32;
            }
        }
        if (small) {
            nt = 32;
            depth = 1;
            const int key = (kind * 64 + 1) * 16 + 1;
            auto it = f.tile_sched.find(key);
            if (it == f.tile_sched.end()) it = f.tile_sched.emplace(key, build_tile_schedule(f, *lv, kind, 32, 1)).first;
            sched = it->second;
            chunk = 4096;
            vt.alloc((size_t)(tiles * f.N * W), ledger);
            mem.alloc((size_t)(tiles * sched->Mpad), ledger);
            oor.alloc(1, ledger);
            return;
        }
        // CTA width: pick the instantiated width that minimises items(nt) · (454 + nt) for this formula's
        // level sizes (a fixed per-item cost plus an issue cost proportional to the width).  With levels cut
        // to the width (list-scheduled EXACT, width-sized BALANCED classes) the measured step time on B200 is
        // flat within 2 % from 768 to 1024 threads (0.590 / 0.603 / 0.603 ms at 768 / 960 / 1024, BALANCED,
        // N = 10 000): the kernel is throughput-bound on padded thread slots, not on the item count.
        {
            int forced = 0;
            if (const char* e = std::getenv("ODESAT_TILE_NT")) {
                const int v = std::atoi(e);
                for (int c : cand) if (v == c) forced = v;
            }
            double best = 1e300;
            for (int c : cand) {
                if (forced && c != forced) continue;
                if (ragged && c != 128 && c != 512) continue;   // widths k_tile_ragged is instantiated with
                // 704 threads exist for one case: f32 BALANCED where a ring of FOUR stages then fits beside the rows
                // (N = 10 000: 160 KB + 4 x 16.5 KB), which the warp-specialised kernel turns into 2 % (measured
                // in one run: 0.5302 ms/step against 0.5417 at 768 threads with a ring of three; ncu showed the
                // consumers of the 768-thread kernel waiting 1.5 cycles per issued instruction for ring data)
                if (c == 704 && !forced && !(sizeof(T) == 4 && kind == ODESAT_SCHED_BALANCED && ws_env != 0 &&
                                              pick_depth(f.N, 704, (int)(f.M / 704 + 256)) == 4)) continue;
                if (pick_depth(f.N, c, (int)(f.M / c + 256)) < 2 && c != 128) continue;   // its ring would not fit
                auto l = levels_for(c);
                double items = 0;
                for (const auto& b : l->bucket) items += (double)((b.size() + c - 1) / c);
                // width preference measured on B200 at the headline size (BALANCED, ms/step at 512/640/768/1024 threads:
                // f32 0.611/0.596/0.588/0.603, f64 0.656/0.637/0.659/0.666): the widest CTA is not the fastest
                // (with the TMA-fed ring, BALANCED f64 is fastest at 768 threads too: 0.600 ms against 0.629 at 640 with cp.async)
                const double pref = c == 1024 ? (sizeof(T) == 4 ? 1.12 : 1.25) : (c == 768 && sizeof(T) == 8 && kind == ODESAT_SCHED_EXACT ? 1.10 : (c == 704 ? 0.93 : 1.0));
                const double cost = items * (454.0 + c) * pref;
                if (cost < best) { best = cost; nt = c; lv = l; }
            }
            if (!lv) { nt = 128; lv = levels_for(128); }
        }
        int want = 0;
        if (const char* e = std::getenv("ODESAT_TILE_D")) want = std::atoi(e);
        depth = pick_depth(f.N, nt, (int)(f.M / nt + 3 * (int64_t)lv->bucket.size() + 16));
        if (depth < 2) throw Error(ODESAT_EUNSUPPORTED, "variables do not fit in shared memory");
        if (want >= 2 && want <= depth) depth = want;
        if (nt == 704) depth = depth >= 4 ? 4 : 2;   // the only rings this width is instantiated with
        if (ragged) depth = depth >= 6 ? 6 : (depth >= 4 ? 4 : 2);   // rings k_tile_ragged is instantiated with
        const int wide = (kind == ODESAT_SCHED_BALANCED && nt >= 512 && ipl != 1) ? 1 + ipl : 0;
        const int key = ((kind * 64 + nt / 32) * 16 + depth) + 65536 * wide;
        auto it = f.tile_sched.find(key);
        if (it == f.tile_sched.end()) it = f.tile_sched.emplace(key, build_tile_schedule(f, *lv, kind, nt, depth)).first;
        sched = it->second;
        if (smem_bytes(f.N, sched->n_items, nt, depth) > kMaxSmem) throw Error(ODESAT_EUNSUPPORTED, "schedule does not fit in shared memory");
        use_tma = tma_env >= 0 ? tma_env == 1 : (kind == ODESAT_SCHED_BALANCED && nt == 768 && depth % 3 == 0);
        // measured (B200, headline size, ms/step): f32 0.5315 against 0.5444 for the TMA kernel on the same wide levels and
        // 0.5554 for round 1's kernel and levels; f64 0.614 against 0.582 (TMA, wide) — so f64 keeps the TMA kernel
        use_ws = ws_env >= 0 ? ws_env == 1 : (kind == ODESAT_SCHED_BALANCED && ((nt == 768 && depth % 3 == 0) || (nt == 704 && depth == 4)) && sizeof(T) == 4);
        {   // the TMA kernel orders a slot's write-back against its next bulk read with a proxy fence issued at the start
            // of the next NON-EMPTY item: at least one such item must lie between the write and the wrap-around refill.
            // The warp-specialised kernel gates the re-read of a slot on the consumption of the stage's PREVIOUS item,
            // which must be a different, later item than the one that wrote the slot: every stage needs two real items.
            int real = 0;
            for (uint32_t it : sched->items) real += ((it >> 20) & 0x7FFu) != 0u;
            if (real < 2 * depth + 2) use_tma = use_ws = false;
        }
        {
            int dev = 0;
            ODESAT_CUDA(cudaGetDevice(&dev));
            ODESAT_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        }
        work.alloc((size_t)tiles + 1, ledger);
        vt.alloc((size_t)(tiles * f.N * W), ledger);
        mem.alloc((size_t)(tiles * sched->Mpad), ledger);
        oor.alloc(1, ledger);
    }
    void reset_control() override { need_rterm = true; oor_valid = false; }
    void snapshot() override {
        if (!vt_snap.p) { vt_snap.alloc(vt.n, ledger_); mem_snap.alloc(mem.n, ledger_); }
        ODESAT_CUDA(cudaMemcpyAsync(vt_snap.p, vt.p, vt.bytes(), cudaMemcpyDeviceToDevice, stream));
        ODESAT_CUDA(cudaMemcpyAsync(mem_snap.p, mem.p, mem.bytes(), cudaMemcpyDeviceToDevice, stream));
    }
    void restore() override {
        ODESAT_REQUIRE(vt_snap.p != nullptr, "restore without a snapshot");
        ODESAT_CUDA(cudaMemcpyAsync(vt.p, vt_snap.p, vt.bytes(), cudaMemcpyDeviceToDevice, stream));
        ODESAT_CUDA(cudaMemcpyAsync(mem.p, mem_snap.p, mem.bytes(), cudaMemcpyDeviceToDevice, stream));
    }

    void geom(int64_t rows, dim3& grid, dim3& block) const {
        int bx = 1;
        while (bx < 256 && bx < tiles) bx <<= 1;
        const int by = 256 / bx;
        block = dim3(bx, by, 1);
        grid = dim3((unsigned)((rows + by - 1) / by), (unsigned)((tiles + bx - 1) / bx), 1);
    }

    int64_t import_state(const T* v, const T* xs, const T* xl, int64_t Rp) override {
        ODESAT_CUDA(cudaMemsetAsync(oor.p, 0, 4, stream));
        dim3 g, b;
        geom(f.N + sched->Mpad, g, b);
        k_tile_import<T><<<g, b, 0, stream>>>(v, xs, xl, Rp, R, f.N, sched->Mpad, sched->d_perm.p, vt.p, mem.p, tiles, oor.p);
        ODESAT_CUDA(cudaGetLastError());
        need_rterm = true;    // decided on the device: the next launch is the STRICT / fast pair keyed on *oor
        oor_valid = true;
        return 1;
    }
    int64_t export_state(T* v, T* xs, T* xl, int64_t Rp) override {
        dim3 g, b;
        geom(f.N + sched->Mpad, g, b);
        k_tile_export<T><<<g, b, 0, stream>>>(v, xs, xl, Rp, R, f.N, sched->Mpad, sched->d_perm.p, vt.p, mem.p, tiles);
        ODESAT_CUDA(cudaGetLastError());
        return 1;
    }

    bool has_direct() const override { return true; }
    int64_t init_mem(const int8_t* xs0) override {
        dim3 g((unsigned)((sched->Mpad + 255) / 256), (unsigned)tiles), b(256);
        k_tile_init_mem<T><<<g, b, 0, stream>>>(xs0, sched->d_perm.p, mem.p, sched->Mpad, tiles);
        ODESAT_CUDA(cudaGetLastError());
        return 1;
    }
    int64_t import_v(const T* v, int64_t Rp) override {
        ODESAT_CUDA(cudaMemsetAsync(oor.p, 0, 4, stream));
        dim3 g((unsigned)((f.N + 31) / 32), (unsigned)((tiles * W + 31) / 32)), b(32, 8);
        k_tile_import_v<T><<<g, b, 0, stream>>>(v, Rp, R, f.N, vt.p, tiles, oor.p);
        ODESAT_CUDA(cudaGetLastError());
        need_rterm = true;
        oor_valid = true;
        return 1;
    }
    int64_t verify_direct(uint32_t* bad) override {
        if ((size_t)f.N > 48 * 1024 || ragged) return 0;   // k_tile_verify reads packed 3-literal entries
        k_tile_verify<T><<<(unsigned)tiles, 512, (size_t)f.N, stream>>>(vt.p, sched->d_entry.p, sched->d_perm.p, f.N, sched->Mpad, R, bad);
        ODESAT_CUDA(cudaGetLastError());
        return 1;
    }
    int64_t assignment_direct(int64_t r, uint8_t* out_dev) override {
        k_tile_assignment<T><<<(unsigned)((f.N + 255) / 256), 256, 0, stream>>>(vt.p, f.N, r, out_dev);
        ODESAT_CUDA(cudaGetLastError());
        return 1;
    }

    template <int NT, int D, bool STRICT> void launch(const TileArgs<T>& a) {
        const size_t smem = smem_bytes(f.N, sched->n_items, NT, D);
        // persistent CTAs + work queue when the one-CTA-per-tile grid would leave a tail wave worth cutting
        if constexpr (!STRICT && (NT == 512 || NT == 640 || NT == 768)) {
            if (queue_fixed && work.p && tiles <= (1 << 24) && a.nsteps > 1) {
                const int ks = pick_ksub(a.nsteps);
                if (ks < a.nsteps || ksub_env > 0) {
                    static uint64_t attr_q = 0;
                    ensure_max_smem(k_tile_fixed<T, NT, D, false, true, true>, (int)kMaxSmem, attr_q);
                    TileWork wk;
                    wk.counter = work.p; wk.done = work.p + 1; wk.tiles = (int)tiles; wk.ksub = ks;
                    wk.nsub = (a.nsteps + ks - 1) / ks;
                    ODESAT_CUDA(cudaMemsetAsync(work.p, 0, ((size_t)tiles + 1) * sizeof(int), stream));
                    const int64_t grid = std::min<int64_t>(tiles * wk.nsub, num_sms);
                    k_tile_fixed<T, NT, D, false, true, true><<<(unsigned)grid, NT, smem, stream>>>(a, wk);
                    return;
                }
            }
        }
        static uint64_t attr_devs = 0;   // per instantiation: devices on which the attribute is set
        ensure_max_smem(k_tile_fixed<T, NT, D, STRICT, (NT < 1024), false>, (int)kMaxSmem, attr_devs);
        k_tile_fixed<T, NT, D, STRICT, (NT < 1024), false><<<(unsigned)tiles, NT, smem, stream>>>(a, TileWork());
    }
    template <int NT, int D> bool launch_tma(const TileArgs<T>& a) {
        const size_t smem = (size_t)f.N * 16 + (size_t)NT * D * 24 + (size_t)(sched->n_items + 2) * 8 + (size_t)D * 8;
        if (smem > kMaxSmem || sched->n_items % D != 0) return false;
        static uint64_t attr_devs = 0;
        ensure_max_smem(k_tile_fixed_tma<T, NT, D>, (int)kMaxSmem, attr_devs);
        k_tile_fixed_tma<T, NT, D><<<(unsigned)tiles, NT, smem, stream>>>(a);
        return true;
    }
    // Steps per sub-chunk of the work queue: the launch's k steps of `tiles` tiles are dealt to G persistent CTAs in
    // rounds of ksub steps; a round costs ksub steps plus the reload of the tile's rows and the refill of the ring
    // (measured: ≈ 0.25 of a step at the headline size).  Many tiles per SM → one sub-chunk (the whole launch).
    bool followed_ = false;
    void set_followed(bool f) override { followed_ = f; }
    int pick_ksub(int k) const {
        if (ksub_env > 0) return std::min(ksub_env, k);
        // measured (B200, one call of 8 sub-batches of 512 replicas, 20 steps): 12.04 ms with sub-chunks of 5 steps, 11.73 ms
        // without — the next sub-batch's persistent CTAs start on the SMs this kernel's last round leaves idle
        if (followed_) return k;
        const int64_t G = std::max(1, num_sms);
        int best = k;
        double best_cost = 1e300;
        for (int ks = k; ks >= 1; --ks) {
            const int64_t nsub = (k + ks - 1) / ks;
            const double cost = (double)((tiles * nsub + G - 1) / G) * ((double)ks + 0.25);
            if (cost < best_cost * 0.995) { best_cost = cost; best = ks; }
        }
        return best;
    }
    int ws_early = [] { const char* e = std::getenv("ODESAT_TILE_WS_EARLY"); return e ? std::atoi(e) : 0; }();
    template <int NT, int D> bool launch_ws(const TileArgs<T>& a) {
        const size_t smem = (size_t)f.N * 16 + (size_t)NT * D * 24 + (size_t)(sched->n_items + 2) * 8 + (size_t)D * 16 + 16;
        if (smem > kMaxSmem || sched->n_items % D != 0 || !work.p || tiles > (1 << 24)) return false;
        static uint64_t attr_devs = 0;
        ensure_max_smem(k_tile_ws<T, NT, D>, (int)kMaxSmem, attr_devs);
        TileWork wk;
        wk.counter = work.p;
        wk.done = work.p + 1;
        wk.tiles = (int)tiles;
        wk.ksub = pick_ksub(a.nsteps);
        wk.nsub = (a.nsteps + wk.ksub - 1) / wk.ksub;
        wk.early = ws_early;
        ODESAT_CUDA(cudaMemsetAsync(work.p, 0, ((size_t)tiles + 1) * sizeof(int), stream));
        const int64_t grid = std::min<int64_t>(tiles * wk.nsub, num_sms);
        k_tile_ws<T, NT, D><<<(unsigned)grid, NT + 32, smem, stream>>>(a, wk);
        return true;
    }
    template <int NT> void launch_d(const TileArgs<T>& a, bool strict) {
        if (strict) { launch<NT, 2, true>(a); return; }
        if (use_ws) {
            if constexpr (NT == 768 || NT == 512) {
                if (depth % 3 == 0 ? launch_ws<NT, 3>(a) : launch_ws<NT, 2>(a)) return;
            }
        }
        if (use_tma) {
            if constexpr (NT == 768 || NT == 512 || NT == 640) {
                if (depth % 3 == 0 ? launch_tma<NT, 3>(a) : launch_tma<NT, 2>(a)) return;
            } else if constexpr (NT == 1024) {
                if (launch_tma<NT, 2>(a)) return;
            }
        }   // ring of 2 divides every schedule padding
        switch (depth) {
            case 2: launch<NT, 2, false>(a); break;
            case 3: launch<NT, 3, false>(a); break;
            case 4: launch<NT, 4, false>(a); break;
            case 5: launch<NT, 5, false>(a); break;
            default: launch<NT, 6, false>(a); break;
        }
    }
    template <bool STRICT> void launch_small(const TileArgs<T>& a) {
        static uint64_t attr_devs = 0;   // per instantiation: devices on which the attribute is set
        ensure_max_smem(k_tile_small<T, STRICT>, (int)kMaxSmem, attr_devs);
        k_tile_small<T, STRICT><<<(unsigned)tiles, 32, smem_small(f.N, sched->n_items), stream>>>(a);
    }
    // 704 threads: the width whose point is a ring of FOUR beside the rows of 10 000 variables (672 / 736 / 640 with the same
    // ring and 576 with a ring of five were measured slower: DESIGN.md §5) — warp-specialised kernel, else the per-thread ring
    template <int NT> void launch_r4(const TileArgs<T>& a, bool strict) {
        if (strict) launch<NT, 2, true>(a);
        else if (!(use_ws && depth == 4 && launch_ws<NT, 4>(a))) {
            if (depth >= 4) launch<NT, 4, false>(a); else launch<NT, 2, false>(a);
        }
    }
    template <int NT, int D, bool STRICT> void launch_ragged(const TileArgs<T>& a) {
        const size_t smem = smem_bytes(f.N, sched->n_items, NT, D);
        if (sched->n_group > 0) {
            static uint64_t attr_g = 0;
            ensure_max_smem(k_tile_ragged<T, NT, D, STRICT, true>, (int)kMaxSmem, attr_g);
            k_tile_ragged<T, NT, D, STRICT, true><<<(unsigned)tiles, NT, smem, stream>>>(a, sched->d_aux.p);
        } else {
            static uint64_t attr_devs = 0;
            ensure_max_smem(k_tile_ragged<T, NT, D, STRICT, false>, (int)kMaxSmem, attr_devs);
            k_tile_ragged<T, NT, D, STRICT, false><<<(unsigned)tiles, NT, smem, stream>>>(a, sched->d_aux.p);
        }
    }
    template <int NT> void launch_ragged_nt(const TileArgs<T>& a, bool strict) {
        if (strict) launch_ragged<NT, 2, true>(a);
        else if (depth == 6) launch_ragged<NT, 6, false>(a);
        else if (depth == 4) launch_ragged<NT, 4, false>(a);
        else launch_ragged<NT, 2, false>(a);
    }
    void launch_nt(const TileArgs<T>& a, bool strict) {
        if (ragged) { if (nt == 128) launch_ragged_nt<128>(a, strict); else launch_ragged_nt<512>(a, strict); return; }
        if (small) { if (strict) launch_small<true>(a); else launch_small<false>(a); return; }
        if (nt == 128) launch_d<128>(a, strict);
        else if (nt == 512) launch_d<512>(a, strict);
        else if (nt == 640) launch_d<640>(a, strict);
        else if (nt == 768) launch_d<768>(a, strict);
        else if (nt == 704) launch_r4<704>(a, strict);
        else launch_d<1024>(a, strict);
    }

    // ---- adaptive steps (tile_adaptive.cuh) ----------------------------------------------------------------
    DevBuf<T> vfull, cmb;   // scratch of the adaptive kernel, allocated on first use
    static size_t smem_adaptive(int64_t N, int n_items, int nt, int d) { return (size_t)N * 16 + (size_t)nt * d * 32 + (size_t)(n_items + 2) * 8; }
    // ring depth of the adaptive kernel (32-byte stages): 3 when the schedule was padded for it and it fits, else 2
    // (the kernel pads the item list to whole rings itself, so any schedule will do; ODESAT_TILE_AD=2/3 overrides)
    int adaptive_depth() const {
        if (small || (ragged && sched->n_group > 0)) return 0;   // group clauses (EXACT, 4..32 literals): fixed steps only
        static const int env = [] { const char* e = std::getenv("ODESAT_TILE_AD"); return e ? std::atoi(e) : 0; }();
        if (env != 2 && smem_adaptive(f.N, sched->n_items, nt, 3) <= kMaxSmem) return 3;
        return smem_adaptive(f.N, sched->n_items, nt, 2) <= kMaxSmem ? 2 : 0;
    }
    bool has_adaptive() const override { return adaptive_depth() != 0; }
    template <int NT, int D, bool STRICT> void launch_adaptive(const TileAdaptArgs<T>& a) {
        if constexpr (NT == 128 || NT == 512) {   // the widths of ragged formulas
            if (ragged) {
                static uint64_t attr_r = 0;
                ensure_max_smem(k_tile_adaptive<T, NT, D, STRICT, true>, (int)kMaxSmem, attr_r);
                k_tile_adaptive<T, NT, D, STRICT, true><<<(unsigned)tiles, NT, smem_adaptive(f.N, sched->n_items, NT, D), stream>>>(a);
                return;
            }
        }
        static uint64_t attr_devs = 0;
        ensure_max_smem(k_tile_adaptive<T, NT, D, STRICT>, (int)kMaxSmem, attr_devs);
        k_tile_adaptive<T, NT, D, STRICT><<<(unsigned)tiles, NT, smem_adaptive(f.N, sched->n_items, NT, D), stream>>>(a);
    }
    template <int NT> void launch_adaptive_nt(const TileAdaptArgs<T>& a, bool strict, int d) {
        if (strict) launch_adaptive<NT, 2, true>(a);
        else if (d == 3) launch_adaptive<NT, 3, false>(a);
        else launch_adaptive<NT, 2, false>(a);
    }
    // warp-specialised persistent form (tile_adaptive_ws.cuh): f32 BALANCED at the widths of k_tile_ws.  ODESAT_TILE_ADWS=0/1.
    int adws_env = [] { const char* e = std::getenv("ODESAT_TILE_ADWS"); return e ? (e[0] == '1' ? 1 : 0) : -1; }();
    template <int NT, int D> bool launch_adaptive_ws(const TileAdaptArgs<T>& a) {
        if constexpr (sizeof(T) == 4) {
            const size_t smem = (size_t)f.N * 16 + (size_t)NT * D * 32 + (size_t)(sched->n_items + 2) * 8 + (size_t)D * 16 + 16;
            if (smem > kMaxSmem || !work.p || tiles > (1 << 24)) return false;
            int real = 0;
            for (uint32_t it : sched->items) real += ((it >> 20) & 0x7FFu) != 0u;
            if (real < 2 * D + 2) return false;   // every stage must hold two real items (see k_tile_ws)
            static uint64_t attr_devs = 0;
            ensure_max_smem(k_tile_adaptive_ws<NT, D>, (int)kMaxSmem, attr_devs);
            TileWork wk;
            wk.counter = work.p;
            wk.done = work.p + 1;
            wk.tiles = (int)tiles;
            wk.ksub = pick_ksub(a.t.nsteps);
            wk.nsub = (a.t.nsteps + wk.ksub - 1) / wk.ksub;
            ODESAT_CUDA(cudaMemsetAsync(work.p, 0, ((size_t)tiles + 1) * sizeof(int), stream));
            const int64_t grid = std::min<int64_t>(tiles * wk.nsub, num_sms);
            k_tile_adaptive_ws<NT, D><<<(unsigned)grid, NT + 32, smem, stream>>>(a, wk);
            return true;
        } else {
            return false;
        }
    }
    bool try_adaptive_ws(const TileAdaptArgs<T>& a) {
        const bool want = adws_env >= 0 ? adws_env == 1 : true;
        if (!want || ragged || sizeof(T) != 4 || sched->kind != ODESAT_SCHED_BALANCED) return false;
        if (nt == 704) return launch_adaptive_ws<704, 3>(a) || launch_adaptive_ws<704, 2>(a);
        if (nt == 768) return launch_adaptive_ws<768, 2>(a);
        return false;
    }
    void launch_adaptive_any(const TileAdaptArgs<T>& a, bool strict, int d) {
        if (!strict && try_adaptive_ws(a)) return;
        if (nt == 128) launch_adaptive_nt<128>(a, strict, d);
        else if (nt == 512) launch_adaptive_nt<512>(a, strict, d);
        else if (nt == 640) launch_adaptive_nt<640>(a, strict, d);
        else if (nt == 704) launch_adaptive_nt<704>(a, strict, d);
        else if (nt == 768) launch_adaptive_nt<768>(a, strict, d);
        else launch_adaptive_nt<1024>(a, strict, d);
    }
    int64_t run_adaptive(T tol, T zeta, int64_t n, int32_t* solved, int64_t step0, T* dt_arr) override {
        const int d = adaptive_depth();
        if (d == 0) throw Error(ODESAT_EUNSUPPORTED, "adaptive steps: the tile kernel's ring does not fit beside this formula's rows (or one-warp tiles); use the gather engine");
        if (!vfull.p) { vfull.alloc(vt.n, ledger_); cmb.alloc((size_t)(tiles * sched->Mpad * W) + 8, ledger_); }   // + slack: bulk copies round an item up to 16 bytes
        int64_t launches = 0;
        const bool zeta_ok = std::isfinite((double)zeta);
        for (int64_t done = 0; done < n;) {
            const int64_t k = std::min<int64_t>(chunk, n - done);
            TileAdaptArgs<T> a;
            a.t.N = f.N; a.t.Mpad = sched->Mpad; a.t.R = R; a.t.n_items = sched->n_items;
            a.t.items = sched->d_items.p; a.t.entry = sched->d_entry.p;
            a.t.vt = vt.p; a.t.mem = mem.p; a.t.solved = solved;
            a.t.zeta = zeta; a.t.xl_max = T(1e4) * T(f.M);
            a.t.step0 = (int32_t)(step0 + done); a.t.nsteps = (int32_t)k; a.t.freeze = 1;
            a.vfull = vfull.p; a.cm = cmb.p; a.dt = dt_arr; a.tol = tol; a.aux = sched->d_aux.p;
            if (!zeta_ok || (need_rterm && !oor_valid)) {   // the literal statements (see run_fixed)
                a.t.nsteps = 1;
                launch_adaptive_any(a, true, 2);
                done += 1;
                if (zeta_ok) need_rterm = false;
                ++launches;
            } else if (need_rterm) {                         // first launch after an import: STRICT / fast pair keyed on *oor
                TileAdaptArgs<T> s1 = a;
                s1.t.nsteps = 1;
                s1.t.oor = oor.p;
                launch_adaptive_any(s1, true, 2);
                a.t.oor = oor.p;
                launch_adaptive_any(a, false, d);
                done += k;
                need_rterm = false;
                launches += 2;
            } else {
                launch_adaptive_any(a, false, d);
                done += k;
                ++launches;
            }
        }
        ODESAT_CUDA(cudaGetLastError());
        return launches;
    }

    int64_t run_fixed(T dt, T zeta, int64_t n, int freeze, int32_t* solved, int64_t step0,
                      const unsigned long long* stop_key = nullptr) override {
        int64_t launches = 0;
        const bool zeta_ok = std::isfinite((double)zeta);
        for (int64_t done = 0; done < n;) {
            const int64_t k = std::min<int64_t>(chunk, n - done);
            TileArgs<T> a;
            a.N = f.N; a.Mpad = sched->Mpad; a.R = R; a.n_items = sched->n_items;
            a.items = sched->d_items.p; a.entry = sched->d_entry.p;
            a.vt = vt.p; a.mem = mem.p; a.solved = solved;
            a.dt = dt; a.zeta = zeta; a.xl_max = T(1e4) * T(f.M);
            a.step0 = (int32_t)(step0 + done); a.nsteps = (int32_t)k; a.freeze = freeze;
            a.stop_key = stop_key;
            if (!zeta_ok || (need_rterm && !oor_valid)) {
                // non-finite zeta (every step) or a state of unknown provenance (first step): the literal statements
                a.nsteps = 1;
                launch_nt(a, true);
                done += 1;
                if (zeta_ok) need_rterm = false;
                ++launches;
            } else if (need_rterm) {
                // first launch after an import: only its first step can see a value outside the fast domain; the
                // device flag decides which kernel of the pair takes it (TileArgs::oor)
                TileArgs<T> s1 = a;
                s1.nsteps = 1;
                s1.oor = oor.p;
                launch_nt(s1, true);
                a.oor = oor.p;
                launch_nt(a, false);
                done += k;
                need_rterm = false;
                launches += 2;
            } else {
                launch_nt(a, false);
                done += k;
                ++launches;
            }
        }
        ODESAT_CUDA(cudaGetLastError());
        return launches;
    }
};

}  // namespace odesat
