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
#include <cmath>
#include <cstdlib>
#include <string>

#include "common.cuh"
#include "formula.hpp"
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
    int nlev = 0;
    const int32_t* goff = nullptr;
    const uint64_t* entry = nullptr;
    T* vt = nullptr;
    typename TileTraits<T>::Mem* mem = nullptr;
    int32_t* solved = nullptr;
    T dt = T(0), zeta = T(0), xl_max = T(0);
    int32_t step0 = 0, nsteps = 0, freeze = 0;
};

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

// RTERM: evaluate the rigidity term of system.rs:73-80 literally.  It is identically ±0 —
// and adding it is a bit-exact no-op — whenever every v lies in [-1, 1] and zeta is finite
// (SURVEY quirk Q1), which holds from the first clamp on; the engine launches the RTERM
// variant only for a chunk whose imported state violates that.
template <typename T, int NT, bool RTERM>
__global__ void __launch_bounds__(NT, 1) k_tile_fixed(const TileArgs<T> a) {
    constexpr int W = TileTraits<T>::W;
    using Row = typename TileTraits<T>::Row;
    using Mem = typename TileTraits<T>::Mem;
    using IO = RowIO<T, W>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Row* rows = reinterpret_cast<Row*>(smem_raw);
    int32_t* s_goff = reinterpret_cast<int32_t*>(smem_raw + (size_t)a.N * sizeof(Row));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    const int64_t tile = blockIdx.x;
    T* vt = a.vt + tile * a.N * W;
    Mem* mem = a.mem + tile * a.Mpad;

    for (int i = tid; i <= a.nlev; i += NT) s_goff[i] = a.goff[i];
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

    const T hi_s = T(1) - Kc<T>::EPSILON;
    for (int s = 0; s < a.nsteps; ++s) {
        bool all_frozen = true;
#pragma unroll
        for (int w = 0; w < W; ++w) all_frozen = all_frozen && frozen[w];
        if (all_frozen) break;
        unsigned unsat_bits = 0;
        // ------------------------------ clause phase -----------------------------------
        // software prefetch: the packed clause and its memories for this warp's NEXT group are
        // requested before the current group is processed (also across level barriers).
        int lev = 0;
        int g = s_goff[0] + warp;
        while (lev < a.nlev && g >= s_goff[lev + 1]) { ++lev; g = s_goff[lev] + warp; }
        uint64_t e_next = 0;
        Mem m_next;
        if (lev < a.nlev) {
            const int64_t c = (int64_t)g * 32 + lane;
            e_next = __ldg(a.entry + c);
            m_next = mem[c];
        }
        for (int L = 0; L < a.nlev; ++L) {
            while (lev == L) {
                const int64_t c = (int64_t)g * 32 + lane;
                const uint64_t e = e_next;
                const Mem mm = m_next;
                // advance to this warp's next group and issue its loads
                g += NW;
                while (lev < a.nlev && g >= s_goff[lev + 1]) { ++lev; if (lev < a.nlev) g = s_goff[lev] + warp; }
                if (lev < a.nlev) {
                    const int64_t cn = (int64_t)g * 32 + lane;
                    e_next = __ldg(a.entry + cn);
                    m_next = mem[cn];
                }
                if (e & TILE_VALID_BIT) {
                    const int i0 = (int)(e & 0xFFFF), i1 = (int)((e >> 16) & 0xFFFF), i2 = (int)((e >> 32) & 0xFFFF);
                    const T q0 = (e >> 48) & 1 ? T(-1) : T(1), q1 = (e >> 49) & 1 ? T(-1) : T(1), q2 = (e >> 50) & 1 ? T(-1) : T(1);
                    T v0[W], v1[W], v2[W], d0[W], d1[W], d2[W], xs[W], xl[W];
                    IO::unpack(rows[i0], v0, d0);
                    IO::unpack(rows[i1], v1, d1);
                    IO::unpack(rows[i2], v2, d2);
                    IO::unpack_mem(mm, xs, xl);
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                        // system.rs:46-57 (q = ±1, so q·v is exact)
                        const T a0 = T(1) - q0 * v0[w], a1 = T(1) - q1 * v1[w], a2 = T(1) - q2 * v2[w];
                        T mn = inf_v<T>(), sm = inf_v<T>();
                        if (a0 < mn) { sm = mn; mn = a0; } else if (a0 < sm) { sm = a0; }
                        if (a1 < mn) { sm = mn; mn = a1; } else if (a1 < sm) { sm = a1; }
                        if (a2 < mn) { sm = mn; mn = a2; } else if (a2 < sm) { sm = a2; }
                        const T cm = T(0.5) * mn;                                        // :60
                        const T wgt = xl[w] * xs[w];
                        T t0 = wgt * ((T(0.5) * q0) * ((a0 != mn) ? mn : sm));           // :64-70, :80
                        T t1 = wgt * ((T(0.5) * q1) * ((a1 != mn) ? mn : sm));
                        T t2 = wgt * ((T(0.5) * q2) * ((a2 != mn) ? mn : sm));
                        if (RTERM) {
                            const T rg = (T(1) + a.zeta * xl[w]) * (T(1) - xs[w]);
                            t0 = t0 + rg * ((cm == a0) ? T(0.5) * (q0 - v0[w]) : T(0));  // :73-77
                            t1 = t1 + rg * ((cm == a1) ? T(0.5) * (q1 - v1[w]) : T(0));
                            t2 = t2 + rg * ((cm == a2) ? T(0.5) * (q2 - v2[w]) : T(0));
                        }
                        d0[w] = d0[w] + t0;
                        d1[w] = d1[w] + t1;
                        d2[w] = d2[w] + t2;
                        const T dxs = (Kc<T>::BETA * (xs[w] + Kc<T>::EPSILON)) * (cm - Kc<T>::GAMMA);   // :84
                        const T dxl = Kc<T>::ALPHA * (cm - Kc<T>::DELTA);                                // :85
                        if (!(cm < Kc<T>::GAMMA)) unsat_bits |= 1u << w;                                 // :88
                        if (!frozen[w]) {
                            xs[w] = euler_clamp(xs[w], dxs, a.dt, Kc<T>::EPSILON, hi_s);                // :94
                            xl[w] = euler_clamp(xl[w], dxl, a.dt, T(1), a.xl_max);                      // :95
                        }
                    }
                    IO::store_dv(rows + i0, d0);
                    IO::store_dv(rows + i1, d1);
                    IO::store_dv(rows + i2, d2);
                    mem[c] = IO::pack_mem(xs, xl);
                }
            }
            __syncthreads();
        }
        // ------------------------------ flags + variable phase ---------------------------
        unsigned any_unsat = 0;   // __syncthreads_or is a boolean OR: one call per replica of the tile
#pragma unroll
        for (int w = 0; w < W; ++w) any_unsat |= (__syncthreads_or((int)((unsat_bits >> w) & 1u)) ? 1u : 0u) << w;
        for (int i = tid; i < a.N; i += NT) {
            T v[W], dv[W];
            IO::unpack(rows[i], v, dv);
#pragma unroll
            for (int w = 0; w < W; ++w) {
                if (!frozen[w]) v[w] = euler_clamp(v[w], dv[w], a.dt, T(-1), T(1));              // :96
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
    for (int i = tid; i < a.N; i += NT) {
        T v[W], dv[W];
        IO::unpack(rows[i], v, dv);
#pragma unroll
        for (int w = 0; w < W; ++w) vt[(int64_t)i * W + w] = v[w];
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
            bad = bad || fabs(x) > T(1);
        }
        if (bad) *out_of_range = 1u;
    } else {
        const int64_t slot = row - N;
        const int m = perm[slot];
        T a[W], b[W];
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const int64_t r = tile * W + w;
            const bool ok = m >= 0 && r < R;
            a[w] = ok ? xs[(int64_t)m * Rp + r] : T(0);
            b[w] = ok ? xl[(int64_t)m * Rp + r] : T(0);
        }
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

template <typename T> struct TileEngine {
    static constexpr int W = TileTraits<T>::W;
    using Mem = typename TileTraits<T>::Mem;
    static constexpr size_t kMaxSmem = 232448 - 1024;   // 227 KB opt-in limit minus static slack

    const odesat_formula& f;
    int64_t R, tiles;
    std::shared_ptr<TileSchedule> sched;
    cudaStream_t stream;
    DevBuf<T> vt;
    DevBuf<Mem> mem;
    DevBuf<unsigned> oor;
    bool need_rterm = true;
    int nt = 1024;
    int chunk = 64;   // Euler steps per launch

    static size_t smem_bytes(int64_t N, int nlev) { return (size_t)N * 16 + (size_t)(nlev + 2) * 4; }

    static bool supports(const odesat_formula& f, int64_t R, std::string* why) {
        auto no = [&](const char* m) { if (why) *why = m; return false; };
        if (R < 1) return no("empty batch");
        if (f.K != 3) return no("needs uniform clause length 3");
        if (!f.distinct_vars) return no("a clause repeats a variable");
        if (f.N > 65535) return no("more than 65535 variables");
        if (smem_bytes(f.N, 4096) > kMaxSmem) return no("variables do not fit in 227 KB of shared memory");
        return true;
    }
    static bool preferred(const odesat_formula&, int64_t R) { return R >= 8; }

    TileEngine(const odesat_formula& f_, int64_t R_, int kind, cudaStream_t st, int64_t* ledger) : f(f_), R(R_), stream(st) {
        tiles = (R + W - 1) / W;
        auto it = f.tile_sched.find(kind);
        if (it == f.tile_sched.end()) it = f.tile_sched.emplace(kind, build_tile_schedule(f, kind)).first;
        sched = it->second;
        if (smem_bytes(f.N, sched->nlev) > kMaxSmem) throw Error(ODESAT_EUNSUPPORTED, "schedule does not fit in shared memory");
        vt.alloc((size_t)(tiles * f.N * W), ledger);
        mem.alloc((size_t)(tiles * sched->Mpad), ledger);
        oor.alloc(1, ledger);
        if (const char* e = std::getenv("ODESAT_TILE_NT")) { const int v = std::atoi(e); if (v == 256 || v == 512 || v == 1024) nt = v; }
        if (const char* e = std::getenv("ODESAT_TILE_CHUNK")) { const int v = std::atoi(e); if (v > 0) chunk = v; }
    }
    void reset_control() { need_rterm = true; }

    void geom(int64_t rows, dim3& grid, dim3& block) const {
        int bx = 1;
        while (bx < 256 && bx < tiles) bx <<= 1;
        const int by = 256 / bx;
        block = dim3(bx, by, 1);
        grid = dim3((unsigned)((rows + by - 1) / by), (unsigned)((tiles + bx - 1) / bx), 1);
    }

    int64_t import_state(const T* v, const T* xs, const T* xl, int64_t Rp) {
        ODESAT_CUDA(cudaMemsetAsync(oor.p, 0, 4, stream));
        dim3 g, b;
        geom(f.N + sched->Mpad, g, b);
        k_tile_import<T><<<g, b, 0, stream>>>(v, xs, xl, Rp, R, f.N, sched->Mpad, sched->d_perm.p, vt.p, mem.p, tiles, oor.p);
        unsigned h = 0;
        ODESAT_CUDA(cudaMemcpyAsync(&h, oor.p, 4, cudaMemcpyDeviceToHost, stream));
        ODESAT_CUDA(cudaStreamSynchronize(stream));
        ODESAT_CUDA(cudaGetLastError());
        need_rterm = h != 0;
        return 1;
    }
    int64_t export_state(T* v, T* xs, T* xl, int64_t Rp) {
        dim3 g, b;
        geom(f.N + sched->Mpad, g, b);
        k_tile_export<T><<<g, b, 0, stream>>>(v, xs, xl, Rp, R, f.N, sched->Mpad, sched->d_perm.p, vt.p, mem.p, tiles);
        ODESAT_CUDA(cudaGetLastError());
        return 1;
    }

    template <int NT, bool RTERM> void launch(const TileArgs<T>& a) {
        const size_t smem = smem_bytes(f.N, sched->nlev);
        static bool attr_set = false;   // per instantiation
        if (!attr_set) {
            ODESAT_CUDA(cudaFuncSetAttribute(k_tile_fixed<T, NT, RTERM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
            attr_set = true;
        }
        k_tile_fixed<T, NT, RTERM><<<(unsigned)tiles, NT, smem, stream>>>(a);
    }
    template <bool RTERM> void launch_nt(const TileArgs<T>& a) {
        if (nt == 1024) launch<1024, RTERM>(a);
        else if (nt == 512) launch<512, RTERM>(a);
        else launch<256, RTERM>(a);
    }

    int64_t run_fixed(T dt, T zeta, int64_t n, int freeze, int32_t* solved, int64_t step0) {
        int64_t launches = 0;
        const bool zeta_ok = std::isfinite((double)zeta);
        for (int64_t done = 0; done < n;) {
            const int64_t k = std::min<int64_t>(chunk, n - done);
            TileArgs<T> a;
            a.N = f.N; a.Mpad = sched->Mpad; a.R = R; a.nlev = sched->nlev;
            a.goff = sched->d_goff.p; a.entry = sched->d_entry.p;
            a.vt = vt.p; a.mem = mem.p; a.solved = solved;
            a.dt = dt; a.zeta = zeta; a.xl_max = T(1e4) * T(f.M);
            a.step0 = (int32_t)(step0 + done); a.nsteps = (int32_t)k; a.freeze = freeze;
            if (need_rterm || !zeta_ok) {
                // only the first step can see |v| > 1; run it alone with the literal rigidity term
                a.nsteps = 1;
                launch_nt<true>(a);
                done += 1;
                if (zeta_ok) need_rterm = false;
            } else {
                launch_nt<false>(a);
                done += k;
            }
            ++launches;
        }
        ODESAT_CUDA(cudaGetLastError());
        return launches;
    }
};

}  // namespace odesat
