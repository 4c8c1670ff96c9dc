// tile_cluster.cuh — the tile engine for formulas whose variables do not fit in ONE SM's shared
// memory: a thread-block CLUSTER of CL CTAs (CL = 1, 2 or 4) owns one replica, the {v, dv} rows are
// distributed round-robin over the CTAs' shared memories (row i lives in CTA i mod CL), and a clause
// gathers / accumulates its three rows through distributed shared memory (ld/st.shared::cluster).
//
//   f32: 8-byte rows  → up to ≈ 27 000 variables per CTA: N = 50 000 (BASELINE configs[4]) needs CL = 2
//   f64: 16-byte rows → ≈ 13 500 per CTA:                 N = 50 000 needs CL = 4
//
// Same algorithm as tile_engine.cuh (level-scheduled clause streaming, cp.async ring, in-place
// {xs, xl} stream, per-replica freeze), with the cluster as one wide CTA: an item is up to CL·NT
// consecutive clause slots of a level, CTA r takes slots [r·NT, (r+1)·NT) of it.  Levels are
// separated by barrier.cluster (release / acquire) split into arrive — right after the level's
// last dv store — and wait — right before the next level's first row load — so the bookkeeping
// and prefetch instructions in between overlap the barrier latency.
//
// One replica per cluster (W = 1): the 16-byte row of two f32 replicas would halve the variables
// a CTA can hold and double the distributed-shared-memory traffic per clause, which is the scarce
// resource here (≈ 20 B/clk per SM against 128 B/clk for local shared memory).
#pragma once
#include "tile_engine.cuh"

namespace odesat {

template <typename T> struct Pair2;
template <> struct Pair2<float> { using type = float2; };
template <> struct Pair2<double> { using type = double2; };

template <typename T> struct CTileArgs {
    int64_t N = 0, Mpad = 0, R = 0;
    int n_items = 0;
    const uint2* items = nullptr;      // [n_items] {slot base, count | last-of-level << 31}, count ≤ CL·NT
    const uint64_t* entry = nullptr;   // [Mpad] packed clauses (3 × 16-bit variable, sign bits)
    T* vt = nullptr;                                   // [R][N]
    typename Pair2<T>::type* mem = nullptr;            // [R][Mpad] {xs, xl} per clause slot
    int32_t* solved = nullptr;
    T dt = T(0), zeta = T(0), xl_max = T(0);
    int32_t step0 = 0, nsteps = 0, freeze = 0;
    const unsigned long long* stop_key = nullptr;   // see TileArgs
    const unsigned* oor = nullptr;                  // see TileArgs
};

// ---- distributed shared memory primitives -------------------------------------------------------
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned mapa_shared(unsigned addr, unsigned rank) {
    unsigned r;
    asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void ldc_row(unsigned addr, float& v, float& dv) {
    asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v), "=f"(dv) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldc_row(unsigned addr, double& v, double& dv) {
    asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v), "=d"(dv) : "r"(addr) : "memory");
}
__device__ __forceinline__ void stc_dv(unsigned addr, float dv) {
    asm volatile("st.shared::cluster.f32 [%0+4], %1;" ::"r"(addr), "f"(dv) : "memory");
}
__device__ __forceinline__ void stc_dv(unsigned addr, double dv) {
    asm volatile("st.shared::cluster.f64 [%0+8], %1;" ::"r"(addr), "d"(dv) : "memory");
}
__device__ __forceinline__ void stc_u32(unsigned addr, unsigned x) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(x) : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <typename T> __device__ __forceinline__ void cp_async_pair(void* smem, const void* gmem) {
    if (sizeof(T) == 4) cp_async8(smem, gmem);
    else cp_async16(smem, gmem);
}

// Shared memory of one CTA:  rows[ceil(N / CL)] | ring[D][NT] | items[n_items] | flags[CL]
__host__ __device__ inline size_t ctile_rows_bytes(int64_t N, int CL, size_t pair) {
    return (((size_t)((N + CL - 1) / CL) * pair) + 15) / 16 * 16;
}

template <typename T, int NT, int D, int CL, bool STRICT>
__global__ void __launch_bounds__(NT, 1) k_ctile_fixed(const CTileArgs<T> a) {
    using Pair = typename Pair2<T>::type;
    constexpr int LOGC = CL == 1 ? 0 : (CL == 2 ? 1 : 2);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n_items = a.n_items;
    const int Nloc = (int)((a.N + CL - 1) / CL);
    Pair* rows = reinterpret_cast<Pair*>(smem_raw);
    Pair* ring = reinterpret_cast<Pair*>(smem_raw + ctile_rows_bytes(a.N, CL, sizeof(Pair)));
    uint2* s_items = reinterpret_cast<uint2*>(ring + D * NT);
    volatile unsigned* s_flags = reinterpret_cast<volatile unsigned*>(s_items + n_items);

    const int s_first = launch_first_step<STRICT>(a);           // grid-uniform (same words for every CTA)
    if (s_first >= a.nsteps) return;
    const unsigned tid = threadIdx.x;
    const unsigned rank = CL > 1 ? cluster_ctarank() : 0u;
    const int64_t rep = blockIdx.x / CL;                         // one replica per cluster
    const unsigned lane_slot = rank * NT + tid;                  // this thread's position inside an item
    T* vt = a.vt + rep * a.N;
    Pair* my_mem = a.mem + rep * a.Mpad + lane_slot;             // + slot base
    const uint2* my_entry = reinterpret_cast<const uint2*>(a.entry) + lane_slot;
    Pair* my_cell = ring + tid;                                  // + k·NT
    const unsigned rows_s = (unsigned)__cvta_generic_to_shared(rows);

    for (int i = tid; i < n_items; i += NT) s_items[i] = a.items[i];
    for (int i = tid; i < Nloc; i += NT) {
        const int64_t g = (int64_t)i * CL + rank;
        Pair p;
        p.x = g < a.N ? vt[g] : T(0);
        p.y = T(0);
        rows[i] = p;
    }
    int32_t solved_at = a.solved[rep];
    bool frozen = a.freeze && solved_at >= 0;
    if (CL > 1) { cluster_arrive(); cluster_wait(); }            // every CTA of the cluster is resident and initialised
    else __syncthreads();

#pragma unroll
    for (int k = 0; k < D; ++k) {
        const uint2 it = s_items[k];
        if (lane_slot < (it.y & 0x7FFFFFFFu)) cp_async_pair<T>(my_cell + k * NT, my_mem + it.x);
        cp_async_commit();
    }
    uint2 it_next = s_items[0];
    uint2 e_next = make_uint2(0u, 0u);
    if (lane_slot < (it_next.y & 0x7FFFFFFFu)) e_next = __ldg(my_entry + it_next.x);

    for (int s = s_first; s < a.nsteps; ++s) {
        if (frozen) break;                                       // cluster-uniform: derived from the shared flags
        bool unsat = false;
        bool pending = false;                                    // arrived at a level barrier, not yet waited
        // ------------------------------ clause phase -----------------------------------
        for (int base = 0; base < n_items; base += D) {
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const int i = base + k;
                const uint2 it = it_next;
                const uint2 e = e_next;
                {
                    const int i1 = (i + 1 == n_items) ? 0 : i + 1;
                    it_next = s_items[i1];
                    if (lane_slot < (it_next.y & 0x7FFFFFFFu)) e_next = __ldg(my_entry + it_next.x);
                }
                cp_async_wait<D - 1>();
                const bool mine = lane_slot < (it.y & 0x7FFFFFFFu);
                unsigned addr[3];
                T q[3];
                if (mine) {
                    const unsigned idx[3] = {e.x & 0xFFFFu, e.x >> 16, e.y & 0xFFFFu};
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const unsigned local = rows_s + (idx[j] >> LOGC) * (unsigned)sizeof(Pair);
                        addr[j] = CL > 1 ? mapa_shared(local, idx[j] & (unsigned)(CL - 1)) : local;
                        q[j] = (e.y >> (16 + j)) & 1u ? T(-1) : T(1);
                    }
                }
                if (pending) {                                   // previous level's dv stores are complete cluster-wide
                    if (CL > 1) cluster_wait();
                    else __syncthreads();
                    pending = false;
                }
                if (mine) {
                    const Pair mm = my_cell[k * NT];
                    T v[3], d[3];
                    if (CL > 1) {
                        ldc_row(addr[0], v[0], d[0]);
                        ldc_row(addr[1], v[1], d[1]);
                        ldc_row(addr[2], v[2], d[2]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            const Pair r = rows[(addr[j] - rows_s) / (unsigned)sizeof(Pair)];
                            v[j] = r.x;
                            d[j] = r.y;
                        }
                    }
                    T xs = mm.x, xl = mm.y;
                    clause_math<T, STRICT>(v, d, q, xs, xl, false, unsat, a.dt, a.zeta, a.xl_max);
                    if (CL > 1) {
                        stc_dv(addr[0], d[0]);
                        stc_dv(addr[1], d[1]);
                        stc_dv(addr[2], d[2]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 3; ++j) rows[(addr[j] - rows_s) / (unsigned)sizeof(Pair)].y = d[j];
                    }
                    Pair o;
                    o.x = xs;
                    o.y = xl;
                    __stcg(my_mem + it.x, o);
                }
                {
                    int nx = i + D;
                    if (nx >= n_items) nx -= n_items;
                    const uint2 itn = s_items[nx];
                    if (lane_slot < (itn.y & 0x7FFFFFFFu)) cp_async_pair<T>(my_cell + k * NT, my_mem + itn.x);
                    cp_async_commit();
                }
                if ((int)it.y < 0) {                             // last item of a level: cluster-uniform
                    if (CL > 1) cluster_arrive();
                    pending = true;
                }
            }
        }
        if (pending) {
            if (CL > 1) cluster_wait();
            else __syncthreads();
        }
        // ------------------------------ flags + variable phase ---------------------------
        unsigned any_unsat = __syncthreads_or((int)unsat) ? 1u : 0u;
        if (CL > 1) {
            // every CTA publishes its flag into every CTA's flags[rank]
            const unsigned flags_s = (unsigned)__cvta_generic_to_shared(const_cast<unsigned*>(s_flags));
            if (tid < CL) stc_u32(mapa_shared(flags_s + rank * 4u, tid), any_unsat);
        }
        for (int i = tid; i < Nloc; i += NT) {                   // this CTA's rows: v ← clamp(v + dt·dv), dv ← 0
            Pair p = rows[i];
            p.x = euler_clamp(p.x, p.y, a.dt, T(-1), T(1));     // :96
            p.y = T(0);
            rows[i] = p;
        }
        if (CL > 1) {
            cluster_arrive();
            cluster_wait();
            any_unsat = 0u;
#pragma unroll
            for (int p = 0; p < CL; ++p) any_unsat |= s_flags[p];
        } else {
            __syncthreads();
        }
        if (!any_unsat) {
            // the pre-update state of this step was all-satisfied (system.rs:149-153); the update above still happened
            if (solved_at < 0) {
                solved_at = a.step0 + s;
                if (tid == 0 && rank == 0) a.solved[rep] = solved_at;
            }
            if (a.freeze) frozen = true;
        }
    }
    cp_async_wait<0>();
    for (int i = tid; i < Nloc; i += NT) {
        const int64_t g = (int64_t)i * CL + rank;
        if (g < a.N) vt[g] = rows[i].x;
    }
}

// canonical replica-major [row][Rp]  ↔  vt[R][N], mem[R][Mpad]{xs, xl}
template <typename T>
__global__ void k_ctile_import(const T* __restrict__ v, const T* __restrict__ xs, const T* __restrict__ xl, int64_t Rp, int64_t R,
                               int64_t N, int64_t Mpad, const int32_t* __restrict__ perm, T* __restrict__ vt,
                               typename Pair2<T>::type* __restrict__ mem, unsigned* __restrict__ out_of_range) {
    const int64_t rep = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t row = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    if (rep >= R || row >= N + Mpad) return;
    if (row < N) {
        const T x = v[row * Rp + rep];
        vt[rep * N + row] = x;
        if (!(fabs(x) <= T(1))) *out_of_range = 1u;
    } else {
        const int64_t slot = row - N;
        const int m = perm[slot];
        typename Pair2<T>::type p;
        p.x = m >= 0 ? xs[(int64_t)m * Rp + rep] : T(0);
        p.y = m >= 0 ? xl[(int64_t)m * Rp + rep] : T(0);
        if (!mem_in_fast_domain(p.x) || !mem_in_fast_domain(p.y)) *out_of_range = 1u;
        mem[rep * Mpad + slot] = p;
    }
}
template <typename T>
__global__ void k_ctile_export(T* __restrict__ v, T* __restrict__ xs, T* __restrict__ xl, int64_t Rp, int64_t R, int64_t N,
                               int64_t Mpad, const int32_t* __restrict__ perm, const T* __restrict__ vt,
                               const typename Pair2<T>::type* __restrict__ mem) {
    const int64_t rep = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t row = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    if (rep >= R || row >= N + Mpad) return;
    if (row < N) {
        v[row * Rp + rep] = vt[rep * N + row];
    } else {
        const int64_t slot = row - N;
        const int m = perm[slot];
        if (m < 0) return;
        const typename Pair2<T>::type p = mem[rep * Mpad + slot];
        xs[(int64_t)m * Rp + rep] = p.x;
        xl[(int64_t)m * Rp + rep] = p.y;
    }
}

template <typename T> struct ClusterTileEngine final : TileBase<T> {
    using Pair = typename Pair2<T>::type;
    static constexpr size_t kMaxSmem = 232448 - 1024;

    const odesat_formula& f;
    int64_t R;
    std::shared_ptr<TileSchedule> sched;
    cudaStream_t stream;
    DevBuf<T> vt, vt_snap;
    DevBuf<Pair> mem, mem_snap;
    DevBuf<unsigned> oor;
    bool need_rterm = true, oor_valid = false;   // see TileEngine
    int64_t* ledger_ = nullptr;
    int cl = 1, nt = 1024, depth = 2, chunk = 64;

    static size_t smem_bytes(int64_t N, int n_items, int cl, int nt, int depth) {
        return ctile_rows_bytes(N, cl, sizeof(Pair)) + (size_t)nt * depth * sizeof(Pair) + (size_t)n_items * 8 + 16;
    }
    // smallest cluster whose CTAs hold their share of the rows beside a ring of depth >= 2
    static int pick_cluster(int64_t N, int nt, int n_items_guess) {
        for (int c : {1, 2, 4})
            if (smem_bytes(N, n_items_guess, c, nt, 2) <= kMaxSmem) return c;
        return 0;
    }
    static int forced_cluster() {
        if (const char* e = std::getenv("ODESAT_TILE_CLUSTER")) {
            const int v = std::atoi(e);
            if (v == 1 || v == 2 || v == 4) return v;
        }
        return 0;
    }
    // cluster size the constructor would pick (0 = does not fit)
    static int natural_cluster(const odesat_formula& f) {
        for (int c_nt : {1024, 512}) {
            const int c = pick_cluster(f.N, c_nt, (int)(f.M / c_nt + 4 * f.max_degree + 64));
            if (c) return c;
        }
        return 0;
    }
    static bool supports(const odesat_formula& f, int64_t R, std::string* why) {
        auto no = [&](const char* m) { if (why) *why = m; return false; };
        if (R < 1) return no("empty batch");
        if (f.K != 3 || f.M < 1) return no("needs uniform clause length 3");
        if (!f.distinct_vars) return no("a clause repeats a variable");
        if (f.N > 65535) return no("more than 65535 variables");
        if (pick_cluster(f.N, 512, (int)(f.M / 512 + 4 * f.max_degree + 64)) == 0)
            return no("variables do not fit in the shared memory of a 4-CTA cluster");
        return true;
    }

    ClusterTileEngine(const odesat_formula& f_, int64_t R_, int kind, cudaStream_t st, int64_t* ledger) : f(f_), R(R_), stream(st), ledger_(ledger) {
        if (const char* e = std::getenv("ODESAT_TILE_CHUNK")) { const int v = std::atoi(e); if (v > 0) chunk = v; }
        const int force_cl = forced_cluster();
        int want_nt = 0, want_d = 0;
        if (const char* e = std::getenv("ODESAT_TILE_NT")) want_nt = std::atoi(e);
        if (const char* e = std::getenv("ODESAT_TILE_D")) want_d = std::atoi(e);
        // widest CTA first: a level should be a few full items
        for (int c_nt : {1024, 512}) {
            if (want_nt == 512 || want_nt == 1024) { if (c_nt != want_nt) continue; }
            const int guess = (int)(f.M / c_nt + 4 * f.max_degree + 64);
            int c = pick_cluster(f.N, c_nt, guess);
            if (force_cl) c = smem_bytes(f.N, guess, force_cl, c_nt, 2) <= kMaxSmem ? force_cl : 0;
            if (c == 0) continue;
            cl = c;
            nt = c_nt;
            break;
        }
        if (pick_cluster(f.N, nt, 8) == 0 && !force_cl) throw Error(ODESAT_EUNSUPPORTED, "variables do not fit in a 4-CTA cluster");
        const int width = cl * nt;
        // BALANCED levels of about three items (fewer cluster barriers); EXACT levels are list-scheduled with
        // the item width as the cap
        const bool exact = kind == ODESAT_SCHED_EXACT;
        const int lkey = exact ? 2 * width : 1 + 16 * (width / 512);
        auto lv = f.tile_levels.find(lkey);
        if (lv == f.tile_levels.end()) lv = f.tile_levels.emplace(lkey, build_tile_levels(f, kind, exact ? width : 3 * width, width)).first;
        for (depth = 4; depth >= 2; --depth) {
            const int guess = (int)(f.M / width + 2 * (int64_t)lv->second->bucket.size() + 16);
            if (smem_bytes(f.N, guess, cl, nt, depth) <= kMaxSmem) break;
        }
        if (depth < 2) throw Error(ODESAT_EUNSUPPORTED, "variables do not fit in shared memory");
        if (want_d >= 2 && want_d <= depth) depth = want_d;
        const int key = 1 << 24 | ((kind * 64 + width / 32) * 16 + depth);
        auto it = f.tile_sched.find(key);
        if (it == f.tile_sched.end()) it = f.tile_sched.emplace(key, build_tile_schedule(f, *lv->second, kind, width, depth, /*upload=*/true, /*wide=*/true)).first;
        sched = it->second;
        if (smem_bytes(f.N, sched->n_items, cl, nt, depth) > kMaxSmem) throw Error(ODESAT_EUNSUPPORTED, "schedule does not fit in shared memory");
        vt.alloc((size_t)(R * f.N), ledger);
        mem.alloc((size_t)(R * sched->Mpad), ledger);
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
        while (bx < 256 && bx < R) bx <<= 1;
        const int by = 256 / bx;
        block = dim3(bx, by, 1);
        grid = dim3((unsigned)((rows + by - 1) / by), (unsigned)((R + bx - 1) / bx), 1);
    }
    int64_t import_state(const T* v, const T* xs, const T* xl, int64_t Rp) override {
        ODESAT_CUDA(cudaMemsetAsync(oor.p, 0, 4, stream));
        dim3 g, b;
        geom(f.N + sched->Mpad, g, b);
        k_ctile_import<T><<<g, b, 0, stream>>>(v, xs, xl, Rp, R, f.N, sched->Mpad, sched->d_perm.p, vt.p, mem.p, oor.p);
        ODESAT_CUDA(cudaGetLastError());
        need_rterm = true;
        oor_valid = true;
        return 1;
    }
    int64_t export_state(T* v, T* xs, T* xl, int64_t Rp) override {
        dim3 g, b;
        geom(f.N + sched->Mpad, g, b);
        k_ctile_export<T><<<g, b, 0, stream>>>(v, xs, xl, Rp, R, f.N, sched->Mpad, sched->d_perm.p, vt.p, mem.p);
        ODESAT_CUDA(cudaGetLastError());
        return 1;
    }

    template <int NT, int D, int CL, bool STRICT> void launch(const CTileArgs<T>& a) {
        const size_t smem = smem_bytes(f.N, sched->n_items, CL, NT, D);
        auto kern = k_ctile_fixed<T, NT, D, CL, STRICT>;
        static uint64_t attr_devs = 0;   // per instantiation: devices on which the attribute is set
        ensure_max_smem(kern, (int)kMaxSmem, attr_devs);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(R * CL), 1, 1);
        cfg.blockDim = dim3(NT, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        ODESAT_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    }
    template <int NT, int CL> void launch_d(const CTileArgs<T>& a, bool strict) {
        if (strict) { launch<NT, 2, CL, true>(a); return; }
        switch (depth) {
            case 2: launch<NT, 2, CL, false>(a); break;
            case 3: launch<NT, 3, CL, false>(a); break;
            default: launch<NT, 4, CL, false>(a); break;
        }
    }
    template <int NT> void launch_c(const CTileArgs<T>& a, bool strict) {
        if (cl == 1) launch_d<NT, 1>(a, strict);
        else if (cl == 2) launch_d<NT, 2>(a, strict);
        else launch_d<NT, 4>(a, strict);
    }

    int64_t run_fixed(T dt, T zeta, int64_t n, int freeze, int32_t* solved, int64_t step0,
                      const unsigned long long* stop_key = nullptr) override {
        int64_t launches = 0;
        const bool zeta_ok = std::isfinite((double)zeta);
        auto go = [&](const CTileArgs<T>& a, bool strict) {
            if (nt == 512) launch_c<512>(a, strict);
            else launch_c<1024>(a, strict);
            ++launches;
        };
        for (int64_t done = 0; done < n;) {
            const int64_t k = std::min<int64_t>(chunk, n - done);
            CTileArgs<T> a;
            a.N = f.N; a.Mpad = sched->Mpad; a.R = R; a.n_items = sched->n_items;
            a.items = sched->d_items2.p; a.entry = sched->d_entry.p;
            a.vt = vt.p; a.mem = mem.p; a.solved = solved;
            a.dt = dt; a.zeta = zeta; a.xl_max = T(1e4) * T(f.M);
            a.step0 = (int32_t)(step0 + done); a.nsteps = (int32_t)k; a.freeze = freeze;
            a.stop_key = stop_key;
            if (!zeta_ok || (need_rterm && !oor_valid)) {   // the literal statements, one step per launch
                a.nsteps = 1;
                go(a, true);
                done += 1;
                if (zeta_ok) need_rterm = false;
            } else if (need_rterm) {                        // first launch after an import: STRICT / fast pair keyed on *oor
                CTileArgs<T> s1 = a;
                s1.nsteps = 1;
                s1.oor = oor.p;
                go(s1, true);
                a.oor = oor.p;
                go(a, false);
                done += k;
                need_rterm = false;
            } else {
                go(a, false);
                done += k;
            }
        }
        ODESAT_CUDA(cudaGetLastError());
        return launches;
    }
};

}  // namespace odesat
