// slab_engine.cuh — the SLAB engine: fixed-step throughput path for formulas whose variables do NOT fit in the shared
// memory of an SM (N = 50 000 of BASELINE configs[4]; the tile engine stops at ≈ 13 000 / 27 000 variables).
//
// Why not the general (gather) engine: it evaluates the RHS as two launches over the whole batch — clause phase
// (writes one contribution per literal and replica), variable phase (adds them in the reference's order) — so the
// contributions, 12 B per clause-eval written and 12 B read, and most of the 12 B of v gathers go to HBM: ncu counts
// 22.7 GB per step against 7.8 GB algorithmic, and the engine runs AT the DRAM roofline of those bytes (3.6 ms/step
// at N = 50 000 x 2 048 replicas).  Walking the batch in slabs with one launch pair per slab keeps them in L2 but
// loses more to launch / drain / ramp bubbles than it gains (measured in round 1).
//
// This engine keeps the same deterministic two-phase RHS (system.rs:25-91 → contributions → ordered sums, so dv is
// BIT-IDENTICAL to the reference's, like the gather engine's) and changes the execution model:
//   * SLABS of S = 32 bytes of replicas (8 f32 / 4 f64): a row of a slab is exactly one 32-byte L2 sector, read
//     by two lanes as 16-byte vectors.  State is stored slab-major — vt[slab][N][S], mem[slab][M][{xs[S], xl[S]}] —
//     so the once-per-step {xs, xl} stream of a slab is one contiguous 13.6 MB run, and a slab's v rows (1.6 MB) and
//     contributions (20 MB) are small against the 126 MB L2.
//   * ONE PERSISTENT KERNEL per chunk of steps.  Work units — (step, slab, clause rows) and (step, slab, variable
//     rows) — are handed out in a fixed order by an atomic ticket: C(0), C(1), V(0), C(2), V(1), ... per step, i.e.
//     the variable phase of a slab runs one slab behind its clause phase, while the contributions are still in L2.
//     Dependencies (V needs all C units of its slab; C of the next step needs all V units; a contribution buffer —
//     a ring of three — needs the V units of the slab that used it before) are monotonic counters polled with
//     ld.acquire; every dependency points to EARLIER tickets, which are held by running CTAs, so waiting cannot
//     deadlock and is almost never needed.  No launch boundaries, no grid-wide barriers, no host in the loop.
//   * the tile kernel's arithmetic: packed f32x2 (two pairs of replicas per thread), 1 − q·v as an exact FMA, the
//     5-op min / second-min network, the identically-zero rigidity term dropped — valid on the same domain
//     (every v finite in [-1, 1], memories in mem_in_fast_domain, zeta finite); the first step after importing a
//     state outside that domain runs through the STRICT instantiation that executes the reference's statements
//     literally, chosen on the device like in the tile engine.
//   * cache policy: {xs, xl} loads / stores are streaming (evict-first), contribution loads too (read once), so the
//     v rows and the not-yet-consumed contributions are what stays in L2.
// Per-replica flags: clause units raise unsat[replica]; the variable unit that owns row 0 of the slab commits
// solved_step (system.rs:149-153: the update of the flagging step still happens) and clears the word.  Frozen
// replicas integrate with dt = 0 (bit-exact freeze, as in the tile engine).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <string>

#include "common.cuh"
#include "formula.hpp"
#include "tile_engine.cuh"

namespace odesat {

template <typename T> struct SlabTraits;
template <> struct SlabTraits<float> {
    static constexpr int V = 4;    // replicas per thread (16 bytes)
    static constexpr int S = 8;    // replicas per slab (32 bytes, two lanes)
    using Vec = float4;
};
template <> struct SlabTraits<double> {
    static constexpr int V = 2;
    static constexpr int S = 4;
    using Vec = double2;
};

constexpr int SLAB_NT = 256;          // threads per CTA (a container of 8 independent warps)
constexpr int SLAB_MAX_BUFS = 6;      // most contribution buffers the ring may have
constexpr int SLAB_ELL = 16;          // occurrences of a variable kept inline (one 64-byte row of slot indices)

template <typename T> struct SlabArgs {
    int64_t N = 0, M = 0, R = 0;
    int nslab = 0, nCu = 0, nVu = 0;
    int c_passes = 8, v_passes = 8;       // rows per warp unit = 16 x passes (even)
    int lag = 1, bufs = 3;                // V(s) is issued `lag` slabs after C(s); contribution ring of `bufs` > lag buffers
    const uint64_t* entry = nullptr;      // [M] var0 | var1 << 20 | var2 << 40 | signs << 60
    const int32_t* voff = nullptr;        // [N + 1]
    const int32_t* occ_slot = nullptr;    // [L] literal slot of each occurrence, sorted (clause, position)
    const int4* occ16 = nullptr;          // [N][4] the first 16 occurrence slots of every variable, -1 padded
    T* vt = nullptr;                      // [nslab][N][S]
    T* mem = nullptr;                     // [nslab][M][2][S]
    T* contrib = nullptr;                 // [bufs][3 M][S]
    uint32_t* unsat = nullptr;            // [nslab * S]
    int32_t* solved = nullptr;            // [R]
    int* ctr = nullptr;                   // [1 + 2 nslab]: ticket, cdone[nslab], vdone[nslab] (zeroed before the launch)
    T dt = T(0), zeta = T(0), xl_max = T(0);
    int32_t step0 = 0, nsteps = 0, freeze = 0;
    const unsigned long long* stop_key = nullptr;
    const unsigned* oor = nullptr;
};

// 16-byte vector accessors.  Everything mutable is read with ld.global.cg (L2 only): the warps are persistent and the
// same addresses are rewritten by other SMs step after step, so a line left in this SM's L1 would be stale.
// The *_stream forms add an L2 evict-first policy for the {xs, xl} stream (touched once per step).  The contributions
// are NOT streamed: their ring is rewritten in place every three slabs, so as long as the lines stay in L2 — dirty —
// they never travel to HBM at all (a first version read them evict-first: ncu counted 5.2 GB per step of contribution
// write-backs, each line evicted right after its only read).
template <typename Vec> __device__ __forceinline__ Vec ld_cg(const void* p) { return __ldcg(reinterpret_cast<const Vec*>(p)); }
template <typename Vec> __device__ __forceinline__ void st_cg(void* p, const Vec& v) { __stcg(reinterpret_cast<Vec*>(p), v); }
__device__ __forceinline__ unsigned long long slab_policy_evict_first() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ unsigned long long slab_policy_evict_last() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ld_hint(const float* p, unsigned long long pol) {
    float4 r;
    asm volatile("ld.global.cg.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ double2 ld_hint(const double* p, unsigned long long pol) {
    double2 r;
    asm volatile("ld.global.cg.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void st_hint(float* p, const float4& v, unsigned long long pol) {
    asm volatile("st.global.cg.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint(double* p, const double2& v, unsigned long long pol) {
    asm volatile("st.global.cg.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}

template <typename T> struct SlabVecIO;
template <> struct SlabVecIO<float> {
    __device__ static void get(const float4& x, float* o) { o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w; }
    __device__ static float4 put(const float* o) { return make_float4(o[0], o[1], o[2], o[3]); }
};
template <> struct SlabVecIO<double> {
    __device__ static void get(const double2& x, double* o) { o[0] = x.x; o[1] = x.y; }
    __device__ static double2 put(const double* o) { return make_double2(o[0], o[1]); }
};

constexpr int SLAB_MAX_PASSES = 4;    // clause passes per unit (cp.async groups in flight)
extern __shared__ __align__(16) unsigned char slab_smem[];   // [max(5 * SLAB_MAX_PASSES, 16)][SLAB_NT] 16-byte cells
__device__ __forceinline__ void cp_async16_hint(void* smem, const void* gmem, unsigned long long pol) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_wait_upto(int n) {   // at most n groups still pending
    switch (n) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    }
}
__device__ __forceinline__ void slab_wait_ge(const int* ctr, int target) {
    while (ld_acquire_gpu(ctr) < target) { __nanosleep(64); }
}

// The two unit bodies are separate (non-inlined) functions: inlined into the ticket loop, ptxas allocated their
// registers jointly and spilled 600 bytes per thread at the 128-register cap; on their own both fit.
template <typename T, bool STRICT>
__device__ __forceinline__ void slab_clause_unit(const SlabArgs<T>& a, int slab, int u, int step, T* cbuf, int h, int rsub, int RPP, unsigned long long pol_stream) {
    constexpr int V = SlabTraits<T>::V, S = SlabTraits<T>::S;
    using Vec = typename SlabTraits<T>::Vec;
    using IO = SlabVecIO<T>;
    // ---- per-replica control -----------------------------------------------------------------
    const int64_t rep0 = (int64_t)slab * S + h * V;
    bool valid[V], frozen[V];
    T dtw[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
        valid[k] = rep0 + k < a.R;
        const int ss = valid[k] ? __ldcg(a.solved + rep0 + k) : 0;
        // clause units run after the previous step's commit; variable units race with this step's commit of a
        // replica that flags NOW, whose update must still happen (system.rs:149-153): only earlier flags freeze
        frozen[k] = !valid[k] || (a.freeze && ss >= 0 && ss < step);
        dtw[k] = frozen[k] ? T(0) : a.dt;
    }
        // ================================ clause unit =========================================
        const T* vt = a.vt + (int64_t)slab * a.N * S + h * V;
        T* mem = a.mem + (int64_t)slab * a.M * 2 * S + h * V;
        const int64_t m0 = (int64_t)u * (RPP * a.c_passes) + rsub;
        const int CP = a.c_passes;
        // Latency hiding without registers: every state load of ALL passes of the unit — three v rows, xs, xl per pass
        // — is issued up front as a cp.async copy into this thread's own shared-memory cells (no destination
        // registers, no scoreboard), one commit group per pass; the passes are then computed as their groups land.
        // Two memory round trips per unit (clause words, then everything else) instead of one per pass.
        Vec* const cells = reinterpret_cast<Vec*>(slab_smem) + threadIdx.x;           // cell (p, c) at [(p * 5 + c) * SLAB_NT]
        auto load_e = [&](int p) { const int64_t m = m0 + (int64_t)RPP * p; return (p < CP && m < a.M) ? __ldg(a.entry + m) : 0ull; };
        uint64_t ew[SLAB_MAX_PASSES];
#pragma unroll
        for (int p = 0; p < SLAB_MAX_PASSES; ++p) ew[p] = load_e(p);
#pragma unroll
        for (int p = 0; p < SLAB_MAX_PASSES; ++p) {
            const int64_t m = m0 + (int64_t)RPP * p;
            if (p < CP && m < a.M) {
                Vec* c = cells + (p * 5) * SLAB_NT;
                cp_async16(c, vt + (int64_t)(ew[p] & 0xFFFFFu) * S);
                cp_async16(c + SLAB_NT, vt + (int64_t)((ew[p] >> 20) & 0xFFFFFu) * S);
                cp_async16(c + 2 * SLAB_NT, vt + (int64_t)((ew[p] >> 40) & 0xFFFFFu) * S);
                cp_async16_hint(c + 3 * SLAB_NT, mem + m * 2 * S, pol_stream);
                cp_async16_hint(c + 4 * SLAB_NT, mem + m * 2 * S + S, pol_stream);
            }
            cp_async_commit();
        }
        bool unsat[V];
        float mx[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int k = 0; k < V; ++k) unsat[k] = false;
        auto compute = [&](int p, uint64_t ee) {
            const int64_t m = m0 + (int64_t)RPP * p;
            if (p >= CP || m >= a.M) return;
            const unsigned sgn = (unsigned)(ee >> 60);
            const T q[3] = {sgn & 1u ? T(-1) : T(1), sgn & 2u ? T(-1) : T(1), sgn & 4u ? T(-1) : T(1)};
            const Vec* cc = cells + (p * 5) * SLAB_NT;
            const Vec xv0 = cc[0], xv1 = cc[SLAB_NT], xv2 = cc[2 * SLAB_NT], xm0 = cc[3 * SLAB_NT], xm1 = cc[4 * SLAB_NT];
            T v[3][V], xs[V], xl[V], c[3][V];
            IO::get(xv0, v[0]); IO::get(xv1, v[1]); IO::get(xv2, v[2]);
            IO::get(xm0, xs); IO::get(xm1, xl);
            if constexpr (!STRICT && sizeof(T) == 4) {
#pragma unroll
                for (int k = 0; k < V; k += 2) {
                    const float2 v2[3] = {make_float2(v[0][k], v[0][k + 1]), make_float2(v[1][k], v[1][k + 1]), make_float2(v[2][k], v[2][k + 1])};
                    float2 d2[3] = {make_float2(0.0f, 0.0f), make_float2(0.0f, 0.0f), make_float2(0.0f, 0.0f)};
                    float2 xs2 = make_float2(xs[k], xs[k + 1]), xl2 = make_float2(xl[k], xl[k + 1]);
                    float mxk[2] = {mx[k], mx[k + 1]};
                    clause_math_f32x2(v2, d2, q, xs2, xl2, mxk, make_float2(dtw[k], dtw[k + 1]), a.xl_max);
                    mx[k] = mxk[0]; mx[k + 1] = mxk[1];
#pragma unroll
                    for (int j = 0; j < 3; ++j) { c[j][k] = d2[j].x; c[j][k + 1] = d2[j].y; }
                    xs[k] = xs2.x; xs[k + 1] = xs2.y; xl[k] = xl2.x; xl[k + 1] = xl2.y;
                }
            } else {
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    const T vv[3] = {v[0][k], v[1][k], v[2][k]};
                    T dd[3] = {T(0), T(0), T(0)};
                    clause_math<T, STRICT>(vv, dd, q, xs[k], xl[k], frozen[k], unsat[k], dtw[k], a.zeta, a.xl_max);
                    c[0][k] = dd[0]; c[1][k] = dd[1]; c[2][k] = dd[2];
                }
            }
            T* cp = cbuf + (3 * m) * S + h * V;
            st_cg<Vec>(cp, IO::put(c[0]));
            st_cg<Vec>(cp + S, IO::put(c[1]));
            st_cg<Vec>(cp + 2 * S, IO::put(c[2]));
            st_hint(mem + m * 2 * S, IO::put(xs), pol_stream);
            st_hint(mem + m * 2 * S + S, IO::put(xl), pol_stream);
        };
#pragma unroll
        for (int p = 0; p < SLAB_MAX_PASSES; ++p) {
            cp_async_wait_upto(SLAB_MAX_PASSES - 1 - p);        // the groups of passes 0 .. p have landed
            compute(p, ew[p]);
        }
        if constexpr (!STRICT && sizeof(T) == 4) {
#pragma unroll
            for (int k = 0; k < V; ++k) unsat[k] = !(mx[k] < 0.5f);
        }
        // system.rs:88-90: raised only, so test first — every unsatisfied clause of a replica hits the same word
#pragma unroll
        for (int k = 0; k < V; ++k)
            if (unsat[k] && valid[k] && __ldcg(a.unsat + rep0 + k) == 0u) __stcg(a.unsat + rep0 + k, 1u);
}

template <typename T, bool STRICT>
__device__ __forceinline__ void slab_var_unit(const SlabArgs<T>& a, int slab, int u, int step, const T* cbuf, int h, int rsub, int RPP, unsigned long long pol_stream) {
    constexpr int V = SlabTraits<T>::V, S = SlabTraits<T>::S;
    using Vec = typename SlabTraits<T>::Vec;
    using IO = SlabVecIO<T>;
    // ---- per-replica control -----------------------------------------------------------------
    const int64_t rep0 = (int64_t)slab * S + h * V;
    bool valid[V], frozen[V];
    T dtw[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
        valid[k] = rep0 + k < a.R;
        const int ss = valid[k] ? __ldcg(a.solved + rep0 + k) : 0;
        // clause units run after the previous step's commit; variable units race with this step's commit of a
        // replica that flags NOW, whose update must still happen (system.rs:149-153): only earlier flags freeze
        frozen[k] = !valid[k] || (a.freeze && ss >= 0 && ss < step);
        dtw[k] = frozen[k] ? T(0) : a.dt;
    }
        // ================================ variable unit =======================================
        T* vt = a.vt + (int64_t)slab * a.N * S + h * V;
        const T* cb = cbuf + h * V;
        const int64_t i0 = (int64_t)u * (RPP * a.v_passes) + rsub;
        const int VP = a.v_passes;
        if (u == 0 && rsub == 0) {   // row 0 of the slab commits the flags: every clause unit of (step, slab) has finished
#pragma unroll
            for (int k = 0; k < V; ++k)
                if (valid[k]) {
                    const int ss = __ldcg(a.solved + rep0 + k);
                    if (ss < 0 && __ldcg(a.unsat + rep0 + k) == 0u) __stcg(a.solved + rep0 + k, step);
                    __stcg(a.unsat + rep0 + k, 0u);
                }
        }
        // the first 16 contributions of the row go through this thread's shared-memory cells as cp.async copies (all in
        // flight together, no registers); slot indices of the next pass are fetched during this one
        Vec* const cells = reinterpret_cast<Vec*>(slab_smem) + threadIdx.x;           // cell k at [k * SLAB_NT]
        int4 ix[2][4];
        auto load_ix = [&](int p, int b) {
            const int64_t i = i0 + (int64_t)RPP * p;
            if (p < VP && i < a.N) {
#pragma unroll
                for (int k = 0; k < 4; ++k) ix[b][k] = __ldg(a.occ16 + i * 4 + k);
            }
        };
        auto var_pass = [&](int p, int b) {
            const int64_t i = i0 + (int64_t)RPP * p;
            if (p >= VP || i >= a.N) return;
            const int id[16] = {ix[b][0].x, ix[b][0].y, ix[b][0].z, ix[b][0].w, ix[b][1].x, ix[b][1].y, ix[b][1].z, ix[b][1].w,
                                ix[b][2].x, ix[b][2].y, ix[b][2].z, ix[b][2].w, ix[b][3].x, ix[b][3].y, ix[b][3].z, ix[b][3].w};
#pragma unroll
            for (int k = 0; k < 16; ++k)
                if (id[k] >= 0) cp_async16_hint(cells + k * SLAB_NT, cb + (int64_t)id[k] * S, pol_stream);
            cp_async_commit();
            const Vec vi = ld_cg<Vec>(vt + i * S);
            T dv[V];
#pragma unroll
            for (int k = 0; k < V; ++k) dv[k] = T(0);                           // system.rs:33
            cp_async_wait_upto(0);
#pragma unroll
            for (int k = 0; k < 16; ++k)
                if (id[k] >= 0) {                                               // added in the reference's order (:80)
                    T x[V];
                    IO::get(cells[k * SLAB_NT], x);
#pragma unroll
                    for (int w = 0; w < V; ++w) dv[w] = dv[w] + x[w];
                }
            if (id[15] >= 0) {                                                  // more than 16 occurrences: the rest from the CSR
                const int e1 = __ldg(a.voff + i + 1);
                for (int eo = __ldg(a.voff + i) + SLAB_ELL; eo < e1; ++eo) {
                    T x[V];
                    IO::get(ld_hint(cb + (int64_t)__ldg(a.occ_slot + eo) * S, pol_stream), x);
#pragma unroll
                    for (int w = 0; w < V; ++w) dv[w] = dv[w] + x[w];
                }
            }
            T v[V];
            IO::get(vi, v);
#pragma unroll
            for (int k = 0; k < V; ++k) {
                if (STRICT) { if (!frozen[k]) v[k] = euler_clamp(v[k], dv[k], a.dt, T(-1), T(1)); }   // :96
                else v[k] = euler_clamp(v[k], dv[k], dtw[k], T(-1), T(1));
            }
            st_cg<Vec>(vt + i * S, IO::put(v));
        };
        load_ix(0, 0);
#pragma unroll 1
        for (int pp = 0; pp < VP; pp += 2) {
            load_ix(pp + 1, 1);
            var_pass(pp, 0);
            load_ix(pp + 2, 0);
            var_pass(pp + 1, 1);
        }
}

// Work is handed out per WARP (a first version used whole CTAs: every unit boundary drained the CTA through five
// block barriers — ncu: 4.5 stall cycles per issued instruction at barriers on top of 14 at the load scoreboard).
// A warp unit is 16 rows x 8 passes of one slab; lanes 2r and 2r + 1 hold the two 16-byte halves of row r.
// CTAG: a unit is taken by a whole CTA (128 rows per pass, block barriers around the hand-over) instead of by a warp.
template <typename T, bool STRICT, bool CTAG>
__global__ void __launch_bounds__(SLAB_NT, 2) k_slab_fixed(const SlabArgs<T> a) {
    constexpr int V = SlabTraits<T>::V, S = SlabTraits<T>::S;
    constexpr unsigned FULL = 0xFFFFFFFFu;
    using Vec = typename SlabTraits<T>::Vec;
    using IO = SlabVecIO<T>;
    const int s_first = launch_first_step<STRICT>(a);   // block-uniform
    if (s_first >= a.nsteps) return;
    __shared__ int s_ticket;
    const int lane = CTAG ? (int)threadIdx.x : (int)(threadIdx.x & 31);   // index within the group that shares a unit
    const int h = lane & 1;                // which 16-byte half of the slab row
    const int rsub = lane >> 1;            // row within a pass
    constexpr int RPP = CTAG ? SLAB_NT / 2 : 16;
    auto gsync = [&]() { if (CTAG) __syncthreads(); else __syncwarp(); };
    const int n = a.nslab, nC = a.nCu, nV = a.nVu;
    const int64_t per_step = (int64_t)n * (nC + nV);
    const int64_t total = per_step * (a.nsteps - s_first);
    int* const cdone = a.ctr + 1;
    int* const vdone = cdone + n;
    const int64_t L3 = 3 * a.M;
    const unsigned long long pol_stream = slab_policy_evict_first();

    int next_ticket = 0;
    if (lane == 0) next_ticket = atomicAdd(a.ctr, 1);
    for (;;) {
        int64_t tk;
        if (CTAG) {
            __syncthreads();                               // everyone is done with the previous s_ticket
            if (lane == 0) s_ticket = next_ticket;
            __syncthreads();
            tk = s_ticket;
        } else tk = __shfl_sync(FULL, next_ticket, 0);
        if (tk >= total) break;
        if (lane == 0) next_ticket = atomicAdd(a.ctr, 1);   // one unit ahead: its latency hides behind this unit's work
        // ---- decode: step t (local), slab, unit type, unit index --------------------------------
        const int t = (int)(tk / per_step);
        int64_t r = tk - (int64_t)t * per_step;
        // per step: groups g = 0 .. n + lag - 1; group g holds the clause units of slab g (g < n), then the variable
        // units of slab g - lag (g >= lag)
        int slab, u;
        bool is_clause;
        const int lag = a.lag;
        if (r < (int64_t)lag * nC) { slab = (int)(r / nC); u = (int)(r - (int64_t)slab * nC); is_clause = true; }
        else {
            r -= (int64_t)lag * nC;
            const int64_t mid = (int64_t)(n - lag) * (nC + nV);
            if (r < mid) {
                const int g = (int)(r / (nC + nV));
                const int w = (int)(r - (int64_t)g * (nC + nV));
                if (w < nC) { slab = lag + g; is_clause = true; u = w; }
                else { slab = g; is_clause = false; u = w - nC; }
            } else {
                r -= mid;
                const int g = (int)(r / nV);
                slab = n - lag + g; is_clause = false; u = (int)(r - (int64_t)g * nV);
            }
        }
        const int step = a.step0 + s_first + t;
        const int64_t G = (int64_t)t * n + slab;           // slab sequence number → contribution buffer
        T* const cbuf = a.contrib + (G % a.bufs) * L3 * S;
        // ---- dependencies (all on earlier tickets) -----------------------------------------------
        if (lane == 0) {
            if (is_clause) {
                if (t > 0) slab_wait_ge(vdone + slab, t * nV);                       // v of this slab updated by the previous step
                if (G >= a.bufs) {                                                   // the buffer's previous user has been summed
                    const int64_t Gp = G - a.bufs;
                    slab_wait_ge(vdone + (int)(Gp % n), ((int)(Gp / n) + 1) * nV);
                }
            } else {
                slab_wait_ge(cdone + slab, (t + 1) * nC);
            }
        }
        gsync();
        if (is_clause) {
            slab_clause_unit<T, STRICT>(a, slab, u, step, cbuf, h, rsub, RPP, pol_stream);
            __threadfence();
            gsync();
            if (lane == 0) atomicAdd(cdone + slab, 1);
        } else {
            slab_var_unit<T, STRICT>(a, slab, u, step, cbuf, h, rsub, RPP, pol_stream);
            __threadfence();
            gsync();
            if (lane == 0) atomicAdd(vdone + slab, 1);
        }
    }
}

// canonical replica-major [row][Rp]  ↔  slab layouts
template <typename T>
__global__ void k_slab_import(const T* __restrict__ v, const T* __restrict__ xs, const T* __restrict__ xl, int64_t Rp, int64_t R,
                              int64_t N, int64_t M, T* __restrict__ vt, T* __restrict__ mem, int64_t nslab, unsigned* __restrict__ oor) {
    constexpr int S = SlabTraits<T>::S;
    const int64_t rep = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t row = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    if (rep >= nslab * S || row >= N + M) return;
    const int64_t slab = rep / S, k = rep % S;
    const bool in = rep < R;
    if (row < N) {
        const T x = in ? v[row * Rp + rep] : T(0);
        vt[(slab * N + row) * S + k] = x;
        if (!(fabs(x) <= T(1))) *oor = 1u;
    } else {
        const int64_t m = row - N;
        const T a = in ? xs[m * Rp + rep] : T(0), b = in ? xl[m * Rp + rep] : T(0);
        mem[((slab * M + m) * 2) * S + k] = a;
        mem[((slab * M + m) * 2 + 1) * S + k] = b;
        if (!mem_in_fast_domain(a) || !mem_in_fast_domain(b)) *oor = 1u;
    }
}
template <typename T>
__global__ void k_slab_export(T* __restrict__ v, T* __restrict__ xs, T* __restrict__ xl, int64_t Rp, int64_t R, int64_t N, int64_t M,
                              const T* __restrict__ vt, const T* __restrict__ mem) {
    constexpr int S = SlabTraits<T>::S;
    const int64_t rep = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t row = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    if (rep >= R || row >= N + M) return;
    const int64_t slab = rep / S, k = rep % S;
    if (row < N) v[row * Rp + rep] = vt[(slab * N + row) * S + k];
    else {
        const int64_t m = row - N;
        xs[m * Rp + rep] = mem[((slab * M + m) * 2) * S + k];
        xl[m * Rp + rep] = mem[((slab * M + m) * 2 + 1) * S + k];
    }
}

template <typename T> struct SlabEngine final : TileBase<T> {
    static constexpr int S = SlabTraits<T>::S;
    const odesat_formula& f;
    int64_t R, nslab;
    cudaStream_t stream;
    DevBuf<T> vt, mem, contrib, vt_snap, mem_snap;
    DevBuf<uint64_t> entry;
    DevBuf<int4> occ16;
    DevBuf<uint32_t> unsat;
    DevBuf<int> ctr;
    DevBuf<unsigned> oor;
    bool need_rterm = true, oor_valid = false;
    int64_t* ledger_ = nullptr;
    int num_sms = 148;
    int chunk = 32;
    int lag = 1, bufs = 3, c_passes = 8, v_passes = 8;
    bool cta_units = false;  // units per CTA (128 rows per pass; measured 6.0 ms/step) or per warp (16; 4.8 ms/step)

    static bool supports(const odesat_formula& f, int64_t R, std::string* why) {
        auto no = [&](const char* m) { if (why) *why = m; return false; };
        if (R < 1) return no("empty batch");
        if (f.K != 3) return no("needs uniform clause length 3");
        if (f.N < 1 || f.N >= (1 << 20)) return no("needs 1 <= varnum < 2^20");
        if (f.M < 1) return no("no clauses");
        return true;
    }
    // AUTO does not take this engine: measured on B200 (N = 50 000, alpha = 4.25, 2 048 replicas, f32) it needs 4.8 ms per
    // step against 3.56 ms for the gather engine — the HBM traffic is down to 13 GB per step (reads at the algorithmic
    // minimum) but every unit is a chain of dependent L2 round trips (ticket, dependency word, clause words, state,
    // store drain before the completion signal) that 16 warps per SM do not cover.  ODESAT_SLAB=1 or an
    // explicit ODESAT_ENGINE_SLAB selects it.
    static bool preferred(const odesat_formula&, int64_t) {
        const char* e = std::getenv("ODESAT_SLAB");
        return e && e[0] == '1';
    }

    SlabEngine(const odesat_formula& f_, int64_t R_, cudaStream_t st, int64_t* ledger) : f(f_), R(R_), stream(st), ledger_(ledger) {
        nslab = (R + S - 1) / S;
        std::vector<uint64_t> h((size_t)f.M);
        for (int64_t m = 0; m < f.M; ++m) {
            uint64_t e = 0;
            for (int j = 0; j < 3; ++j) {
                const int32_t l = f.h_lits[f.h_off[m] + j];
                const uint64_t var = (uint64_t)((l < 0 ? -l : l) - 1);
                e |= var << (20 * j);
                if (l < 0) e |= 1ull << (60 + j);
            }
            h[m] = e;
        }
        entry.alloc((size_t)f.M, ledger);
        ODESAT_CUDA(cudaMemcpy(entry.p, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
        {   // the first 16 occurrence slots of every variable, in the reference's summation order, -1 padded
            std::vector<int32_t> o((size_t)f.N * SLAB_ELL, -1);
            for (int64_t i = 0; i < f.N; ++i)
                for (int k = 0; k < SLAB_ELL && f.h_voff[i] + k < f.h_voff[i + 1]; ++k) o[(size_t)i * SLAB_ELL + k] = f.h_occ_slot[f.h_voff[i] + k];
            occ16.alloc((size_t)f.N * 4, ledger);
            ODESAT_CUDA(cudaMemcpy(occ16.p, o.data(), o.size() * 4, cudaMemcpyHostToDevice));
        }
        vt.alloc((size_t)(nslab * f.N * S), ledger);
        mem.alloc((size_t)(nslab * f.M * 2 * S), ledger);
        if (const char* e = std::getenv("ODESAT_SLAB_CTA")) cta_units = e[0] != '0';
        c_passes = 4; v_passes = 2;
        if (const char* e = std::getenv("ODESAT_SLAB_LAG")) lag = std::max(1, std::atoi(e));
        if (const char* e = std::getenv("ODESAT_SLAB_BUFS")) bufs = std::atoi(e);
        if (const char* e = std::getenv("ODESAT_SLAB_CPASSES")) c_passes = std::max(2, std::atoi(e) & ~1);
        if (const char* e = std::getenv("ODESAT_SLAB_VPASSES")) v_passes = std::max(2, std::atoi(e) & ~1);
        c_passes = std::min(c_passes, SLAB_MAX_PASSES);
        lag = (int)std::min<int64_t>(lag, nslab);
        bufs = std::min(SLAB_MAX_BUFS, std::max(bufs, lag + 1));
        contrib.alloc((size_t)(bufs * 3 * f.M * S), ledger);
        unsat.alloc((size_t)(nslab * S), ledger);
        ctr.alloc((size_t)(1 + 2 * nslab), ledger);
        oor.alloc(1, ledger);
        ODESAT_CUDA(cudaMemsetAsync(unsat.p, 0, unsat.bytes(), stream));
        int dev = 0;
        ODESAT_CUDA(cudaGetDevice(&dev));
        ODESAT_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        if (const char* e = std::getenv("ODESAT_SLAB_CHUNK")) { const int v = std::atoi(e); if (v > 0) chunk = v; }
    }
    void reset_control() override {
        need_rterm = true;
        oor_valid = false;
        ODESAT_CUDA(cudaMemsetAsync(unsat.p, 0, unsat.bytes(), stream));
    }
    void snapshot() override {
        if (!vt_snap.p) { vt_snap.alloc(vt.n, ledger_); mem_snap.alloc(mem.n, ledger_); }
        ODESAT_CUDA(cudaMemcpyAsync(vt_snap.p, vt.p, vt.bytes(), cudaMemcpyDeviceToDevice, stream));
        ODESAT_CUDA(cudaMemcpyAsync(mem_snap.p, mem.p, mem.bytes(), cudaMemcpyDeviceToDevice, stream));
    }
    void restore() override {
        ODESAT_REQUIRE(vt_snap.p != nullptr, "restore without a snapshot");
        ODESAT_CUDA(cudaMemcpyAsync(vt.p, vt_snap.p, vt.bytes(), cudaMemcpyDeviceToDevice, stream));
        ODESAT_CUDA(cudaMemcpyAsync(mem.p, mem_snap.p, mem.bytes(), cudaMemcpyDeviceToDevice, stream));
        ODESAT_CUDA(cudaMemsetAsync(unsat.p, 0, unsat.bytes(), stream));
    }
    void geom(dim3& grid, dim3& block) const {
        const int64_t reps = nslab * S;
        int bx = 1;
        while (bx < 256 && bx < reps) bx <<= 1;
        const int by = 256 / bx;
        block = dim3(bx, by, 1);
        grid = dim3((unsigned)((f.N + f.M + by - 1) / by), (unsigned)((reps + bx - 1) / bx), 1);
    }
    int64_t import_state(const T* v, const T* xs, const T* xl, int64_t Rp) override {
        ODESAT_CUDA(cudaMemsetAsync(oor.p, 0, 4, stream));
        dim3 g, b;
        geom(g, b);
        k_slab_import<T><<<g, b, 0, stream>>>(v, xs, xl, Rp, R, f.N, f.M, vt.p, mem.p, nslab, oor.p);
        ODESAT_CUDA(cudaGetLastError());
        need_rterm = true;
        oor_valid = true;
        return 1;
    }
    int64_t export_state(T* v, T* xs, T* xl, int64_t Rp) override {
        dim3 g, b;
        geom(g, b);
        k_slab_export<T><<<g, b, 0, stream>>>(v, xs, xl, Rp, R, f.N, f.M, vt.p, mem.p);
        ODESAT_CUDA(cudaGetLastError());
        return 1;
    }

    static constexpr size_t kSmem = (size_t)(5 * SLAB_MAX_PASSES > 16 ? 5 * SLAB_MAX_PASSES : 16) * SLAB_NT * 16;   // 80 KB
    void launch(const SlabArgs<T>& a, bool strict) {
        ensure_attr();
        ODESAT_CUDA(cudaMemsetAsync(ctr.p, 0, ctr.bytes(), stream));
        const int64_t units = (int64_t)a.nslab * (a.nCu + a.nVu) * a.nsteps;
        const int per_cta = cta_units ? 1 : SLAB_NT / 32;
        const int64_t grid = std::min<int64_t>((units + per_cta - 1) / per_cta, (int64_t)num_sms * 2);
        if (strict) {
            if (cta_units) k_slab_fixed<T, true, true><<<(unsigned)grid, SLAB_NT, kSmem, stream>>>(a);
            else k_slab_fixed<T, true, false><<<(unsigned)grid, SLAB_NT, kSmem, stream>>>(a);
        } else {
            if (cta_units) k_slab_fixed<T, false, true><<<(unsigned)grid, SLAB_NT, kSmem, stream>>>(a);
            else k_slab_fixed<T, false, false><<<(unsigned)grid, SLAB_NT, kSmem, stream>>>(a);
        }
    }

    void ensure_attr() {
        static uint64_t devs[4] = {0, 0, 0, 0};
        ensure_max_smem(k_slab_fixed<T, true, true>, (int)kSmem, devs[0]);
        ensure_max_smem(k_slab_fixed<T, true, false>, (int)kSmem, devs[1]);
        ensure_max_smem(k_slab_fixed<T, false, true>, (int)kSmem, devs[2]);
        ensure_max_smem(k_slab_fixed<T, false, false>, (int)kSmem, devs[3]);
    }
    int64_t run_fixed(T dt, T zeta, int64_t n, int freeze, int32_t* solved, int64_t step0,
                      const unsigned long long* stop_key = nullptr) override {
        int64_t launches = 0;
        const bool zeta_ok = std::isfinite((double)zeta);
        for (int64_t done = 0; done < n;) {
            const int64_t k = std::min<int64_t>(chunk, n - done);
            SlabArgs<T> a;
            a.N = f.N; a.M = f.M; a.R = R;
            a.nslab = (int)nslab;
            a.c_passes = c_passes; a.v_passes = v_passes; a.lag = lag; a.bufs = bufs;
            const int rpp = cta_units ? SLAB_NT / 2 : 16;
            a.nCu = (int)((f.M + rpp * c_passes - 1) / (rpp * c_passes));
            a.nVu = (int)((f.N + rpp * v_passes - 1) / (rpp * v_passes));
            a.entry = entry.p; a.voff = f.dev.voff; a.occ_slot = f.dev.occ_slot; a.occ16 = occ16.p;
            a.vt = vt.p; a.mem = mem.p; a.contrib = contrib.p; a.unsat = unsat.p; a.solved = solved; a.ctr = ctr.p;
            a.dt = dt; a.zeta = zeta; a.xl_max = T(1e4) * T(f.M);
            a.step0 = (int32_t)(step0 + done); a.nsteps = (int32_t)k; a.freeze = freeze;
            a.stop_key = stop_key;
            if (!zeta_ok || (need_rterm && !oor_valid)) {
                a.nsteps = 1;
                launch(a, true);
                done += 1;
                if (zeta_ok) need_rterm = false;
                ++launches;
            } else if (need_rterm) {
                SlabArgs<T> s1 = a;
                s1.nsteps = 1;
                s1.oor = oor.p;
                launch(s1, true);
                a.oor = oor.p;
                launch(a, false);
                done += k;
                need_rterm = false;
                launches += 2;
            } else {
                launch(a, false);
                done += k;
                ++launches;
            }
        }
        ODESAT_CUDA(cudaGetLastError());
        return launches;
    }
};

}  // namespace odesat
