// batch.cuh — device-resident replica batch: owns the state in HBM, drives the kernels.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <memory>

#include "formula.hpp"
#include "kernels_gather.cuh"
#include "kernels_small.cuh"
#include "tile_cluster.cuh"
#include "tile_engine.cuh"
#include "slab_engine.cuh"

#ifndef ODESAT_CLAUSE_HALF_VEC
#define ODESAT_CLAUSE_HALF_VEC 0   // measured on B200: half-width clause vectors lose (9.8 vs 8.4 ms at N = 50k)
#endif

namespace odesat {

// internal engine code (not part of the C ABI): AUTO restricted to engines that integrate adaptive steps
constexpr int ENGINE_AUTO_ADAPTIVE = 16;

struct BatchBase {
    const odesat_formula* f = nullptr;
    int64_t R = 0, Rp = 0;
    int precision = ODESAT_F64, engine = ODESAT_ENGINE_GATHER, schedule = ODESAT_SCHED_EXACT;
    int64_t launches = 0, dev_bytes = 0;
    int64_t step = 0;   // Euler steps issued since init/upload
    int device = 0;     // CUDA device that owns the buffers and the stream
    // another shard's work on the SAME device is enqueued right behind this shard's (sub-batches of one call): its kernel
    // fills this kernel's tail wave, so the tile engine need not cut the launch into sub-chunks to balance the SMs
    bool followed = false;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // Early-exit words, KEY_SLOTS slots of {min key, unflagged replicas}: device copy (read by the next chunk's
    // kernels as their stop key), pinned host mirror (read by the polling host one chunk late), one event per slot.
    static constexpr int KEY_SLOTS = 3;                  // 0/1: chunk parity of the step loop; 2: synchronous queries
    unsigned long long* d_key = nullptr;
    unsigned long long* h_key = nullptr;
    cudaEvent_t ev_key[KEY_SLOTS] = {nullptr, nullptr, nullptr};
    virtual ~BatchBase() {
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        for (cudaEvent_t e : ev_key) if (e) cudaEventDestroy(e);
        if (d_key) cudaFree(d_key);
        if (h_key) cudaFreeHost(h_key);
        if (stream) cudaStreamDestroy(stream);
    }
    void sync() {
        ODESAT_CUDA(cudaStreamSynchronize(stream));
        ODESAT_CUDA(cudaGetLastError());
    }
    const unsigned long long* key_dev(int slot) const { return d_key + 2 * slot; }
    // enqueue: key / unflagged count of the current flags → device slot → pinned host slot; never blocks
    virtual void post_key(int slot, int64_t replica_offset) = 0;
    // the same reduction into caller-owned device words (no host mirror)
    virtual void post_key_to(int64_t replica_offset, unsigned long long* out_dev) = 0;
    // block until the slot's words are on the host
    void wait_key(int slot, int64_t* key, int64_t* unflagged) {
        ODESAT_CUDA(cudaEventSynchronize(ev_key[slot]));
        if (key) *key = (int64_t)h_key[2 * slot];
        if (unflagged) *unflagged = (int64_t)h_key[2 * slot + 1];
    }
    // step loop pieces that only ENQUEUE work (the synchronous run_* below add timing and a stream sync)
    virtual void run_fixed_async(double dt, double zeta, int64_t n, int freeze, const unsigned long long* stop_key) = 0;
    virtual void run_adaptive_async(double tol, double zeta, int64_t n) = 0;
    // device-to-device copy of state + flags + step counter, and back (lock-step `inter`, see drive())
    virtual void snapshot() = 0;
    virtual void restore() = 0;
    virtual void reset() = 0;   // flags, step counter, dt back to the state of a fresh batch
    // finalize = false: an upload() of the remaining arrays follows immediately (saves one layout conversion)
    virtual void init(uint64_t seed, int64_t replica_offset, bool gen_v, bool gen_xs, bool gen_xl, bool finalize = true) = 0;
    // enqueue only; the host buffers must stay valid until the next sync()
    virtual void upload(const void* v, const void* xs, const void* xl, bool reset) = 0;
    virtual void download(void* v, void* xs, void* xl) = 0;
    virtual void run_fixed(double dt, double zeta, int64_t n, int freeze, float* ms) = 0;
    virtual void run_adaptive(double tol, double zeta, int64_t n, float* ms) = 0;
    // adaptive `simulate_inter` with the reference's shared dt (system.rs:312-349): sequential over replicas;
    // runs until some replica flags or max_steps (< 0: unbounded) outer steps are done; returns the outer steps run
    virtual int64_t run_inter_adaptive(double tol, double zeta, int64_t max_steps) = 0;
    virtual void status(int64_t* solved, int64_t* steps) = 0;
    virtual int64_t first_key() = 0;
    virtual int preferred_chunk() const { return 32; }   // Euler steps between early-exit polls when the caller does not say
    virtual void verify(uint8_t* out) = 0;
    virtual void assignment(int64_t r, uint8_t* out) = 0;
    // flags + exact verification of every replica with ONE synchronisation: enqueue (verify kernel, copies into pinned
    // host staging), then — after the other shards have enqueued theirs — collect
    virtual void results_enqueue() = 0;
    virtual void results_collect(int64_t* solved_out, uint8_t* verified_out) = 0;
    virtual void get_dt(double* out) = 0;
    virtual void set_dt(const double* in) = 0;
    // single-state helpers (R == 1 mirrors of odesat::system)
    virtual void derivatives(double zeta, void* dv, void* dxs, void* dxl, int* allsat) = 0;
    virtual void update_state_with(const void* dv, const void* dxs, const void* dxl, double dt) = 0;
    virtual double max_error_vs(const void* bv, const void* bxs, const void* bxl) = 0;
};

template <typename T> struct StateBuf {
    DevBuf<T> v, xs, xl;
    void alloc(int64_t N, int64_t M, int64_t Rp, int64_t* ledger) {
        v.alloc((size_t)(N * Rp), ledger);
        xs.alloc((size_t)(M * Rp), ledger);
        xl.alloc((size_t)(M * Rp), ledger);
    }
    bool allocated() const { return v.p != nullptr; }
};

template <typename T> struct BatchImpl final : BatchBase {
    using U = typename ErrBits<T>::U;
    // ---- gather engine state ----
    StateBuf<T> S[2], H, Fb;
    int cur = 0;
    DevBuf<int32_t> solved;
    DevBuf<uint32_t> unsat;     // [Rp]
    DevBuf<T> contrib;          // [L][Rp] per-literal contributions (allocated on first use)
    DevBuf<U> err;
    DevBuf<T> dtv;
    DevBuf<T> staging;
    DevBuf<uint8_t> small8;
    DevBuf<double> dscratch;
    DevBuf<uint32_t> bad_buf;   // [R] verification result (some clause falsified)
    int32_t* h_solved = nullptr;   // pinned staging of results_enqueue / results_collect
    uint32_t* h_bad = nullptr;
    StateBuf<T> Snap;           // snapshot of the gather-engine state
    DevBuf<int32_t> solved_snap;
    int64_t step_snap = 0;
    bool v_in_range_snap = false;
    // the formula as this batch's kernels see it: the formula's own device view, or — single large instances on the gather
    // engine — its sorted view (formula.hpp): clause rows stored by smallest variable, xs / xl rows permuted on the way in / out
    FormulaDev fd;
    const int32_t* cperm = nullptr;   // storage position → original clause (sorted view), device
    // replica batches on the gather engine: measured on B200 at N = 50 000 x 2 048 replicas (configs[4]'s share), f32:
    // 3.568 -> 3.474 ms/step with the sorted view (one of the three v rows of a clause is then shared with its neighbours
    // in L2); default from N = 30 000 (smaller formulas run on the tile engine).  ODESAT_GATHER_SORT_BATCH_N overrides.
    static int64_t sort_min_n_batch() {
        const char* e = std::getenv("ODESAT_GATHER_SORT_BATCH_N");
        return e ? std::atoll(e) : 30000;
    }
    // ---- tile engine state ----
    std::unique_ptr<TileBase<T>> tile;   // TileEngine (one CTA per tile) or ClusterTileEngine (one cluster per replica)

    BatchImpl(const odesat_formula* f_, int64_t R_, int engine_, int schedule_) {
        // AUTO for a run of adaptive steps: like AUTO, but only engines that have an adaptive kernel
        const bool need_adaptive = engine_ == ENGINE_AUTO_ADAPTIVE;
        if (need_adaptive) engine_ = ODESAT_ENGINE_AUTO;
        f = f_;
        R = R_;
        Rp = R >= 32 ? pad32(R) : R;
        precision = sizeof(T) == 4 ? ODESAT_F32 : ODESAT_F64;
        schedule = schedule_;
        ODESAT_CUDA(cudaGetDevice(&device));
        ODESAT_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        ODESAT_CUDA(cudaEventCreate(&ev0));
        ODESAT_CUDA(cudaEventCreate(&ev1));
        for (cudaEvent_t& e : ev_key) ODESAT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ODESAT_CUDA(cudaMalloc((void**)&d_key, 2 * KEY_SLOTS * sizeof(unsigned long long)));
        ODESAT_CUDA(cudaHostAlloc((void**)&h_key, 2 * KEY_SLOTS * sizeof(unsigned long long), cudaHostAllocDefault));
        dev_bytes += 2 * KEY_SLOTS * (int64_t)sizeof(unsigned long long);
        std::string why, why_c;
        const bool tile_ok = TileEngine<T>::supports(*f, R, &why);
        const bool ctile_ok = ClusterTileEngine<T>::supports(*f, R, &why_c);
        const bool force_c = ClusterTileEngine<T>::forced_cluster() != 0;
        if (engine_ == ODESAT_ENGINE_TILE && !tile_ok && !ctile_ok)
            throw Error(ODESAT_EUNSUPPORTED, "tile engine cannot run this formula: " + why + "; " + why_c);
        // Measured on B200 (N = 10 000, 43 000 clauses of which 3 000 have 5 or 8 literals, 4 096 replicas, f32): loop clauses
        // (tile_ragged.cuh) make the EXACT schedule's 94 one-item levels wait for one slow warp each — 2.54 ms/step against
        // 1.66 on the gather engine — while BALANCED's wide levels absorb them (1.27).  One- and two-literal clauses cost
        // nothing extra (0.66 / 0.62 ms against 1.60).
        const bool loopy_exact = f->n_loopy > 0 && schedule_ == ODESAT_SCHED_EXACT;
        const bool want_tile = engine_ == ODESAT_ENGINE_TILE || (engine_ == ODESAT_ENGINE_AUTO && TileEngine<T>::preferred(*f, R) && !loopy_exact);
        // Measured on B200 (N = 50 000, 2 048 replicas, f32): with 2 CTAs per replica the scattered 8-byte
        // distributed-shared-memory gathers go through the generic LD/ST path at one lane per cycle and the
        // step takes 12.3 ms against 7.3 ms for the general engine — so AUTO takes the cluster engine only
        // when the rows fit in ONE CTA (no remote traffic); wider clusters on explicit request.
        const bool c_auto = ctile_ok && ClusterTileEngine<T>::natural_cluster(*f) == 1;
        const bool use_ctile = want_tile && ctile_ok && !tile_ok ? (force_c || engine_ == ODESAT_ENGINE_TILE || c_auto)
                                                                 : (want_tile && ctile_ok && force_c);
        const bool use_tile = want_tile && tile_ok && !use_ctile;
        engine = (use_tile || use_ctile) ? ODESAT_ENGINE_TILE : ODESAT_ENGINE_GATHER;
        // SLAB: formulas too large for shared memory, batches with enough replicas to fill the GPU with 32-byte slabs
        std::string why_s;
        const bool slab_ok = SlabEngine<T>::supports(*f, R, &why_s);
        if (engine_ == ODESAT_ENGINE_SLAB && !slab_ok) throw Error(ODESAT_EUNSUPPORTED, "slab engine cannot run this formula: " + why_s);
        if (engine_ == ODESAT_ENGINE_SLAB || (engine_ == ODESAT_ENGINE_AUTO && !need_adaptive && engine == ODESAT_ENGINE_GATHER && slab_ok && SlabEngine<T>::preferred(*f, R)))
            engine = ODESAT_ENGINE_SLAB;
        const int64_t N = f->N, M = f->M;
        solved.alloc((size_t)std::max<int64_t>(Rp, 1), &dev_bytes);
        dtv.alloc((size_t)std::max<int64_t>(Rp, 1), &dev_bytes);
        if (engine == ODESAT_ENGINE_TILE) {
            // supports() works from estimates; the compiled schedule can still turn out too large for shared memory.
            // Under AUTO fall back one engine at a time instead of failing the call.
            try {
                if (use_ctile) tile.reset(new ClusterTileEngine<T>(*f, R, schedule, stream, &dev_bytes));
                else tile.reset(new TileEngine<T>(*f, R, schedule, stream, &dev_bytes));
            } catch (const Error& e) {
                if (e.code != ODESAT_EUNSUPPORTED || engine_ != ODESAT_ENGINE_AUTO) throw;
                tile.reset();
                if (!use_ctile && c_auto) {
                    try { tile.reset(new ClusterTileEngine<T>(*f, R, schedule, stream, &dev_bytes)); }
                    catch (const Error& e2) { if (e2.code != ODESAT_EUNSUPPORTED) throw; tile.reset(); }
                }
                if (!tile) engine = ODESAT_ENGINE_GATHER;
            }
            if (need_adaptive && tile && !tile->has_adaptive()) { tile.reset(); engine = ODESAT_ENGINE_GATHER; }
        }
        if (engine == ODESAT_ENGINE_SLAB) tile.reset(new SlabEngine<T>(*f, R, stream, &dev_bytes));
        fd = f->dev;
        {   // Measured on B200 (N = 1 M, alpha = 4.2, R = 1, adaptive): see DESIGN §4.  ODESAT_GATHER_SORT=0/1 overrides.
            const char* e = std::getenv("ODESAT_GATHER_SORT");
            const bool want = e ? e[0] != '0' : (R == 1 ? f->N >= 100000 : f->N >= sort_min_n_batch());
            if (!tile && R >= 1 && want && f->M > 0 && !small_ok(true) && !small_ok(false)) {
                const auto& sv = f->sorted_view();
                fd = sv.dev;
                cperm = sv.cperm.p;
            }
        }
        if (!tile) {
            S[0].alloc(N, M, Rp, &dev_bytes);   // S[1] (derivatives / adaptive ping-pong) on first use
            pick_slab();
            unsat.alloc((size_t)std::max<int64_t>(Rp, 1), &dev_bytes);
            err.alloc((size_t)std::max<int64_t>(Rp, 1), &dev_bytes);
        }
        reset_control();
    }

    // -------------------------------------------------------------------------------------
    void geom(int64_t rows, dim3& grid, dim3& block) const {
        int bx = 1;
        while (bx < 256 && bx < R) bx <<= 1;
        const int by = 256 / bx;
        block = dim3(bx, by, 1);
        grid = dim3((unsigned)std::max<int64_t>((rows + by - 1) / by, 1), (unsigned)std::max<int64_t>((R + bx - 1) / bx, 1), 1);
    }

    void reset_control() {
        step = 0;
        cur = 0;
        ODESAT_CUDA(cudaMemsetAsync(solved.p, 0xFF, solved.bytes(), stream));   // -1
        if (unsat.p) ODESAT_CUDA(cudaMemsetAsync(unsat.p, 0, unsat.bytes(), stream));
        if (err.p) ODESAT_CUDA(cudaMemsetAsync(err.p, 0, err.bytes(), stream));
        {
            const int64_t n = std::max<int64_t>(Rp, 1);
            k_fill<T><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(dtv.p, n, T(0.01));   // system.rs:205
        }
        if (tile) tile->reset_control();
    }

    void reset() override { reset_control(); }
    void ensure_alt() { if (!S[1].allocated()) S[1].alloc(f->N, f->M, Rp, &dev_bytes); }

    void ensure_staging() {
        const size_t need = (size_t)(std::max(f->N, f->M) * std::max<int64_t>(R, 1));
        if (staging.n < need) staging.alloc(need, &dev_bytes);
    }

    // host [R][X] → device [X][Rp]
    void put(const void* host, T* dst, int64_t X, bool clause_rows = false) {
        if (X == 0 || R == 0) return;
        if (R == 1 && clause_rows && cperm) {   // sorted view: dst[p] = host[cperm[p]]
            ensure_staging();
            ODESAT_CUDA(cudaMemcpyAsync(staging.p, host, (size_t)X * sizeof(T), cudaMemcpyHostToDevice, stream));
            k_permute_rows<T><<<(unsigned)((X + 255) / 256), 256, 0, stream>>>(staging.p, dst, cperm, X, /*gather=*/1);
            ++launches;
            return;
        }
        if (R == 1) {
            ODESAT_CUDA(cudaMemcpyAsync(dst, host, (size_t)X * sizeof(T), cudaMemcpyHostToDevice, stream));
            return;
        }
        ensure_staging();
        ODESAT_CUDA(cudaMemcpyAsync(staging.p, host, (size_t)(R * X) * sizeof(T), cudaMemcpyHostToDevice, stream));
        dim3 g((unsigned)((X + 31) / 32), (unsigned)((R + 31) / 32)), b(32, 8);
        k_transpose_in<T><<<g, b, 0, stream>>>(staging.p, dst, R, X, Rp, clause_rows ? cperm : nullptr);
        ++launches;
    }
    void get(void* host, const T* src, int64_t X, bool clause_rows = false) {
        if (X == 0 || R == 0) return;
        if (R == 1 && clause_rows && cperm) {   // sorted view: host[cperm[p]] = src[p]
            ensure_staging();
            k_permute_rows<T><<<(unsigned)((X + 255) / 256), 256, 0, stream>>>(src, staging.p, cperm, X, /*gather=*/0);
            ++launches;
            ODESAT_CUDA(cudaMemcpyAsync(host, staging.p, (size_t)X * sizeof(T), cudaMemcpyDeviceToHost, stream));
            ODESAT_CUDA(cudaStreamSynchronize(stream));   // staging is reused by the next get()
            return;
        }
        if (R == 1) {
            ODESAT_CUDA(cudaMemcpyAsync(host, src, (size_t)X * sizeof(T), cudaMemcpyDeviceToHost, stream));
            return;
        }
        ensure_staging();
        dim3 g((unsigned)((X + 31) / 32), (unsigned)((R + 31) / 32)), b(32, 8);
        k_transpose_out<T><<<g, b, 0, stream>>>(src, staging.p, R, X, Rp, clause_rows ? cperm : nullptr);
        ++launches;
        ODESAT_CUDA(cudaMemcpyAsync(host, staging.p, (size_t)(R * X) * sizeof(T), cudaMemcpyDeviceToHost, stream));
        ODESAT_CUDA(cudaStreamSynchronize(stream));   // staging is reused by the next get()
    }

    // Canonical replica-major views of the current state (tile engine converts on demand).
    StateBuf<T>& canon() {
        if (tile) {
            if (!S[0].allocated()) S[0].alloc(f->N, f->M, Rp, &dev_bytes);
            return S[0];
        }
        return S[cur];
    }
    bool canon_current = false;   // S[0] mirrors the tile engine's state (skip redundant exports)
    bool canon_ahead = false;     // S[0] holds NEWER data than the tile layout (init without finalize): import before stepping
    void tile_to_canon() {
        if (tile && !canon_current && !canon_ahead) {
            canon();
            launches += tile->export_state(S[0].v.p, S[0].xs.p, S[0].xl.p, Rp);
            canon_current = true;
        }
    }
    void canon_to_tile() {
        if (tile) {
            launches += tile->import_state(S[0].v.p, S[0].xs.p, S[0].xl.p, Rp);
            canon_current = true;
            canon_ahead = false;
        }
    }

    void init(uint64_t seed, int64_t replica_offset, bool gen_v, bool gen_xs, bool gen_xl, bool finalize = true) override {
        StateBuf<T>& s = canon();
        if (tile && tile->has_direct() && gen_xs && gen_xl && !canon_ahead && R > 0) {
            // memories straight into the tile layout; v (if generated) through the canonical scratch and a v-only
            // import — no full-state layout conversion.  When !gen_v the tile keeps its v (or upload(v) follows).
            launches += tile->init_mem(fd.xs0);
            if (gen_v && f->N > 0) {
                dim3 g, b;
                geom(f->N, g, b);
                k_init_state<T><<<g, b, 0, stream>>>(s.v.p, s.xs.p, s.xl.p, fd.xs0, f->N, f->M, R, Rp, seed, replica_offset, 1, 0, 0);
                ++launches;
                launches += tile->import_v(s.v.p, Rp);
            }
            canon_current = false;
            ODESAT_CUDA(cudaGetLastError());
            return;
        }
        if (tile && !(gen_v && gen_xs && gen_xl) && !finalize) {
            // the caller uploads the arrays that are not generated here: nothing of the old state survives
        } else if (tile && !(gen_v && gen_xs && gen_xl)) tile_to_canon();
        dim3 g, b;
        geom(f->N + f->M, g, b);
        if (f->N + f->M > 0 && R > 0) {
            k_init_state<T><<<g, b, 0, stream>>>(s.v.p, s.xs.p, s.xl.p, fd.xs0, f->N, f->M, R, Rp, seed,
                                                  replica_offset, gen_v, gen_xs, gen_xl);
            ++launches;
        }
        ODESAT_CUDA(cudaGetLastError());
        if (gen_v) v_in_range = true;                    // v0 ∈ [-1, 1) by construction
        if (tile && !finalize) { canon_ahead = true; canon_current = false; return; }
        canon_to_tile();
    }

    void upload(const void* v, const void* xs, const void* xl, bool reset) override {
        if (reset) reset_control();
        StateBuf<T>& s = canon();
        if (tile && tile->has_direct() && v && !xs && !xl && !canon_ahead && R > 0) {   // v only: the tile keeps its memories
            put(v, s.v.p, f->N);
            launches += tile->import_v(s.v.p, Rp);
            canon_current = false;
            ODESAT_CUDA(cudaGetLastError());
            return;
        }
        if (tile && !(v && xs && xl)) tile_to_canon();
        if (v) put(v, s.v.p, f->N);
        if (xs) put(xs, s.xs.p, f->M, true);
        if (xl) put(xl, s.xl.p, f->M, true);
        ODESAT_CUDA(cudaGetLastError());
        canon_to_tile();
        if (!tile) check_range();
    }

    void download(void* v, void* xs, void* xl) override {
        tile_to_canon();
        StateBuf<T>& s = canon();
        if (v) get(v, s.v.p, f->N);
        if (xs) get(xs, s.xs.p, f->M, true);
        if (xl) get(xl, s.xl.p, f->M, true);
        ODESAT_CUDA(cudaGetLastError());
        ODESAT_CUDA(cudaStreamSynchronize(stream));
    }

    // ---- slabs (experiment, off by default) -----------------------------------------------------
    // ODESAT_GATHER_SLAB=S walks the batch in slabs of S replicas, `ODESAT_GATHER_SLAB_STEPS` Euler
    // steps per visit, with ONE contribution buffer [L][S] reused by every slab, so that a slab's v
    // rows and contributions stay in the 126 MB L2.  Measured on B200 (N = 50 000, 2 048 replicas,
    // f32): 7.7 / 6.1 / 4.7 / 4.1 / 3.9 ms per step for S = 16 / 32 / 128 / 256 / 512 against 3.74 ms
    // unslabbed — the small launches are instruction- and latency-bound and lose more than the L2
    // residency gains, so the default is one launch pair over the whole batch.
    // the streaming clause kernel may use the tile kernel's shorter arithmetic once every v is known to be finite in
    // [-1, 1] (true after any step: system.rs:96 clamps; checked on upload) and zeta is finite
    bool v_in_range = false;
    DevBuf<unsigned> range_flag;
    // the fast arithmetic needs every v finite in [-1, 1] and every memory in mem_in_fast_domain (tile_engine.cuh)
    void check_range() {
        if (tile || R == 0 || f->N + f->M == 0) { v_in_range = true; return; }
        if (!range_flag.p) range_flag.alloc(1, &dev_bytes);
        ODESAT_CUDA(cudaMemsetAsync(range_flag.p, 0, 4, stream));
        dim3 g, b;
        geom(f->N + f->M, g, b);
        k_check_range<T><<<g, b, 0, stream>>>(S[cur].v.p, S[cur].xs.p, S[cur].xl.p, f->N, f->M, R, Rp, range_flag.p);
        ++launches;
        unsigned h = 1;
        ODESAT_CUDA(cudaMemcpyAsync(&h, range_flag.p, 4, cudaMemcpyDeviceToHost, stream));
        ODESAT_CUDA(cudaStreamSynchronize(stream));
        v_in_range = h == 0;
    }
    int64_t slab = 0;        // 0 = whole batch in one launch pair
    int slab_steps = 1;
    void pick_slab() {
        slab = 0;
        slab_steps = 1;
        const char* e = std::getenv("ODESAT_GATHER_SLAB");
        if (!e) return;
        const int64_t s = std::atol(e);
        if (s <= 0 || s >= Rp || s % (16 / (int64_t)sizeof(T)) != 0 || Rp % (16 / (int64_t)sizeof(T)) != 0) return;
        slab = s;
        slab_steps = 8;
        if (const char* e2 = std::getenv("ODESAT_GATHER_SLAB_STEPS")) { const int v = std::atoi(e2); if (v > 0) slab_steps = v; }
    }

    GatherArgs<T> base_args(double zeta, bool slabbed = false) {
        GatherArgs<T> a;
        a.f = fd;
        a.R = R;
        a.Rp = Rp;
        a.rep0 = 0;
        a.rep1 = R;
        a.cstride = (slabbed && slab > 0) ? slab : Rp;
        a.zeta = (T)zeta;
        a.fast = (v_in_range && std::isfinite(zeta)) ? 1 : 0;
        if (const char* e = std::getenv("ODESAT_GATHER_FAST")) { if (e[0] == '0') a.fast = 0; }
        if (const char* e = std::getenv("ODESAT_GATHER_PACKED")) { if (e[0] == '0' && a.fast) a.fast = 2; }   // 2: scalar fast path (A/B)
        if (const char* e = std::getenv("ODESAT_GATHER_L2HINTS")) a.l2_hints = std::atoi(e);
        a.xl_max = T(1e4) * T(f->M);
        a.solved_step = solved.p;
        a.unsat = unsat.p;
        a.err = err.p;
        const size_t need = (size_t)std::max<int64_t>(f->L * a.cstride, 1);
        if (contrib.n < need) contrib.alloc(need, &dev_bytes);
        a.contrib = contrib.p;
        return a;
    }

    // one RHS evaluation + update = clause phase, then variable phase; 16 bytes of replicas per
    // thread when the rows are 16-byte aligned (R >= 32 ⇒ Rp is a multiple of 32)
    static constexpr int VW = 16 / (int)sizeof(T);
    template <int V> void geom_v(int64_t rows, int64_t reps, int rows_per_thread, dim3& grid, dim3& block, int bx_cap = 256) const {
        const int64_t rv = (reps + V - 1) / V;
        int bx = 1;
        while (bx < bx_cap && bx < rv) bx <<= 1;
        const int by = 256 / bx;
        block = dim3(bx, by, 1);
        const int64_t rpb = (int64_t)by * rows_per_thread;
        grid = dim3((unsigned)std::max<int64_t>((rows + rpb - 1) / rpb, 1), (unsigned)std::max<int64_t>((rv + bx - 1) / bx, 1), 1);
    }
    template <int MODE, int V> void launch_gather_v(GatherArgs<T>& a) {
        dim3 g, b;
        const int64_t reps = a.rep1 - a.rep0;
        // a slab launch has few replicas: one row per thread keeps enough blocks in flight
        a.rows_per_thread = (a.cstride != Rp) ? 1 : 8;
        if (f->M > 0) {
            // the clause phase is latency-bound on registers: half-width vectors double its occupancy
            constexpr int VC = (V > 1 && ODESAT_CLAUSE_HALF_VEC) ? V / 2 : V;
            static const bool stream_on = [] { const char* e = std::getenv("ODESAT_GATHER_STREAM"); return !(e && e[0] == '0'); }();
            if (f->K == 3 && stream_on) {
                // streaming clause phase: RPT rows per thread, all their loads in flight as cp.async copies
                constexpr int CB = VC * (int)sizeof(T);
                constexpr int RPT = CB == 16 ? 2 : 4;
                int cap = 256;
                if (const char* e = std::getenv("ODESAT_GATHER_BX")) { const int v = std::atoi(e); if (v >= 1 && v <= 256) cap = v; }
                geom_v<VC>(f->M, reps, RPT, g, b, cap);
                k_clause_stream<T, MODE, VC, RPT><<<g, b, (size_t)RPT * 5 * 256 * CB, stream>>>(a);
            } else {
                geom_v<VC>(f->M, reps, a.rows_per_thread, g, b);
                if (f->K == 3) k_clause_phase<T, 3, MODE, VC><<<g, b, 0, stream>>>(a);
                else k_clause_phase<T, 0, MODE, VC><<<g, b, 0, stream>>>(a);
            }
            ++launches;
        }
        geom_v<V>(std::max<int64_t>(f->N, 1), reps, a.rows_per_thread, g, b);
        k_var_phase<T, MODE, V><<<g, b, 0, stream>>>(a);
        ++launches;
    }
    template <int MODE> void launch_gather(GatherArgs<T>& a) {
        if (R == 0) return;
        if (Rp % VW == 0 && R >= 32) launch_gather_v<MODE, VW>(a);
        else launch_gather_v<MODE, 1>(a);
    }

    // ---- persistent small-instance kernel (kernels_small.cuh) ------------------------------------
    static constexpr size_t kSmallSmem = 200 * 1024;
    bool small_ok(bool adaptive) const {
        const char* e = std::getenv("ODESAT_SMALL");
        const bool on = !(e && e[0] == '0');
        return on && !tile && R >= 1 && f->N + f->M > 0 &&
               small_smem_bytes(f->N, f->M, f->L, adaptive, sizeof(T)) <= kSmallSmem;
    }
    int preferred_chunk() const override { return small_ok(true) ? 1024 : 32; }
    template <int NT> void launch_small_nt(const SmallArgs<T>& a, size_t smem) {
        static uint64_t attr_devs = 0;   // per instantiation: devices on which the attribute is set
        ensure_max_smem(k_solve_small<T, NT>, (int)kSmallSmem, attr_devs);
        k_solve_small<T, NT><<<(unsigned)R, NT, smem, stream>>>(a);
    }
    void run_small(bool adaptive, double dt, double tol, double zeta, int64_t n, int freeze, const unsigned long long* stop_key) {
        SmallArgs<T> a;
        a.stop_key = stop_key;
        a.f = fd;
        a.R = R; a.Rp = Rp;
        a.v = S[cur].v.p; a.xs = S[cur].xs.p; a.xl = S[cur].xl.p;
        a.dt_arr = dtv.p;
        a.solved_step = solved.p;
        a.dt = (T)dt; a.tol = (T)tol; a.zeta = (T)zeta; a.xl_max = T(1e4) * T(f->M);
        a.freeze = freeze;
        a.adaptive = adaptive ? 1 : 0;
        const size_t smem = small_smem_bytes(f->N, f->M, f->L, adaptive, sizeof(T));
        for (int64_t done = 0; done < n;) {   // int32 step arguments: at most 2^30 steps per launch
            const int64_t k = std::min<int64_t>(n - done, int64_t(1) << 30);
            a.step0 = (int32_t)(step + done);
            a.nsteps = (int32_t)k;
            if (f->M > 1024) launch_small_nt<1024>(a, smem);
            else launch_small_nt<256>(a, smem);
            ++launches;
            done += k;
        }
        ODESAT_CUDA(cudaGetLastError());
    }

    void time_begin(float* ms) { if (ms) ODESAT_CUDA(cudaEventRecord(ev0, stream)); }
    void time_end(float* ms) {
        if (ms) {
            ODESAT_CUDA(cudaEventRecord(ev1, stream));
            ODESAT_CUDA(cudaEventSynchronize(ev1));
            ODESAT_CUDA(cudaEventElapsedTime(ms, ev0, ev1));
        }
        ODESAT_CUDA(cudaStreamSynchronize(stream));
        ODESAT_CUDA(cudaGetLastError());
    }

    void run_fixed(double dt, double zeta, int64_t n, int freeze, float* ms) override {
        time_begin(ms);
        run_fixed_async(dt, zeta, n, freeze, nullptr);
        time_end(ms);
    }
    void run_fixed_async(double dt, double zeta, int64_t n, int freeze, const unsigned long long* stop_key) override {
        ODESAT_REQUIRE(n >= 0, "negative step count");
        ODESAT_REQUIRE(step + n < (int64_t(1) << 31) - 4, "step counter overflow");
        if (tile) {
            if (canon_ahead) canon_to_tile();
            if (n > 0) canon_current = false;
            tile->set_followed(followed);
            launches += tile->run_fixed((T)dt, (T)zeta, n, freeze, solved.p, step, stop_key);
            step += n;
            return;
        }
        if (small_ok(false)) {
            run_small(false, dt, 0.0, zeta, n, freeze, stop_key);
            step += n;
            return;
        }
        auto one_step = [&](int64_t r0, int64_t r1, int64_t at_step) {
            GatherArgs<T> a = base_args(zeta, true);   // in place: each element is read and written by its own thread
            a.rep0 = r0; a.rep1 = r1;
            a.v = S[cur].v.p; a.xs = S[cur].xs.p; a.xl = S[cur].xl.p;
            a.ov = S[cur].v.p; a.oxs = S[cur].xs.p; a.oxl = S[cur].xl.p;
            a.dt = (T)dt;
            a.step = (int32_t)at_step;
            a.freeze = freeze;
            a.stop_key = stop_key;
            launch_gather<G_FIXED>(a);
            v_in_range = true;                           // every stepped v is clamped (system.rs:96); frozen replicas were clamped before
        };
        if (slab > 0) {
            for (int64_t done = 0; done < n; done += slab_steps) {
                const int64_t k = std::min<int64_t>(slab_steps, n - done);
                for (int64_t r0 = 0; r0 < R; r0 += slab)
                    for (int64_t i = 0; i < k; ++i) one_step(r0, std::min<int64_t>(r0 + slab, R), step + done + i);
            }
        } else {
            for (int64_t i = 0; i < n; ++i) one_step(0, R, step + i);
        }
        step += n;
        ODESAT_CUDA(cudaGetLastError());
    }

    void run_adaptive(double tol, double zeta, int64_t n, float* ms) override {
        time_begin(ms);
        run_adaptive_async(tol, zeta, n);
        time_end(ms);
    }
    void run_adaptive_async(double tol, double zeta, int64_t n) override {
        ODESAT_REQUIRE(n >= 0, "negative step count");
        ODESAT_REQUIRE(step + n < (int64_t(1) << 31) - 4, "step counter overflow");
        if (tile) {   // tile_adaptive.cuh (throws ODESAT_EUNSUPPORTED on engines / formulas without an adaptive kernel)
            if (canon_ahead) canon_to_tile();
            if (n > 0) canon_current = false;
            launches += tile->run_adaptive((T)tol, (T)zeta, n, solved.p, step, dtv.p);
            step += n;
            return;
        }
        if (small_ok(true)) {
            run_small(true, 0.0, tol, zeta, n, 1, nullptr);
            step += n;
            return;
        }
        ensure_alt();
        if (!H.allocated()) { H.alloc(f->N, f->M, Rp, &dev_bytes); Fb.alloc(f->N, f->M, Rp, &dev_bytes); }
        for (int64_t i = 0; i < n; ++i) {
            GatherArgs<T> a = base_args(zeta);
            a.dt_arr = dtv.p;
            a.step = (int32_t)step;
            // pass A: k1 on y → y_half (H), y_full (F), flags
            a.v = S[cur].v.p; a.xs = S[cur].xs.p; a.xl = S[cur].xl.p;
            a.ov = H.v.p; a.oxs = H.xs.p; a.oxl = H.xl.p;
            a.fv = Fb.v.p; a.fxs = Fb.xs.p; a.fxl = Fb.xl.p;
            a.yv = S[cur].v.p; a.yxs = S[cur].xs.p; a.yxl = S[cur].xl.p;
            launch_gather<G_ADAPT_A>(a);
            // pass B: k2 on y_half → y_new, error norm
            a.yv = S[cur].v.p; a.yxs = S[cur].xs.p; a.yxl = S[cur].xl.p;
            a.v = H.v.p; a.xs = H.xs.p; a.xl = H.xl.p;
            a.ov = S[1 - cur].v.p; a.oxs = S[1 - cur].xs.p; a.oxl = S[1 - cur].xl.p;
            launch_gather<G_ADAPT_B>(a);
            if (R > 0) {
                k_adapt_c<T><<<(unsigned)((R + 255) / 256), 256, 0, stream>>>(solved.p, unsat.p, err.p, dtv.p, (T)tol, R, (int32_t)step);
                ++launches;
            }
            cur = 1 - cur;
            ++step;
            v_in_range = true;   // stepped replicas are clamped; flagged ones are never evaluated again
        }
        ODESAT_CUDA(cudaGetLastError());
    }

    // ---- snapshot / restore (lock-step `inter`) --------------------------------------------------
    void snapshot() override {
        if (!solved_snap.p) solved_snap.alloc(solved.n, &dev_bytes);
        ODESAT_CUDA(cudaMemcpyAsync(solved_snap.p, solved.p, solved.bytes(), cudaMemcpyDeviceToDevice, stream));
        step_snap = step;
        v_in_range_snap = v_in_range;
        if (tile) {
            if (canon_ahead) canon_to_tile();
            tile->snapshot();
            return;
        }
        if (!Snap.allocated()) Snap.alloc(f->N, f->M, Rp, &dev_bytes);
        ODESAT_CUDA(cudaMemcpyAsync(Snap.v.p, S[cur].v.p, S[cur].v.bytes(), cudaMemcpyDeviceToDevice, stream));
        ODESAT_CUDA(cudaMemcpyAsync(Snap.xs.p, S[cur].xs.p, S[cur].xs.bytes(), cudaMemcpyDeviceToDevice, stream));
        ODESAT_CUDA(cudaMemcpyAsync(Snap.xl.p, S[cur].xl.p, S[cur].xl.bytes(), cudaMemcpyDeviceToDevice, stream));
    }
    void restore() override {
        ODESAT_REQUIRE(solved_snap.p != nullptr, "restore without a snapshot");
        ODESAT_CUDA(cudaMemcpyAsync(solved.p, solved_snap.p, solved.bytes(), cudaMemcpyDeviceToDevice, stream));
        step = step_snap;
        v_in_range = v_in_range_snap;
        if (tile) {
            tile->restore();
            canon_current = false;
            return;
        }
        ODESAT_CUDA(cudaMemcpyAsync(S[cur].v.p, Snap.v.p, S[cur].v.bytes(), cudaMemcpyDeviceToDevice, stream));
        ODESAT_CUDA(cudaMemcpyAsync(S[cur].xs.p, Snap.xs.p, S[cur].xs.bytes(), cudaMemcpyDeviceToDevice, stream));
        ODESAT_CUDA(cudaMemcpyAsync(S[cur].xl.p, Snap.xl.p, S[cur].xl.bytes(), cudaMemcpyDeviceToDevice, stream));
    }

    int64_t run_inter_adaptive(double tol, double zeta, int64_t max_steps) override {
        if (tile) throw Error(ODESAT_EUNSUPPORTED, "adaptive inter runs on the gather engine");
        if (R == 0) return 0;
        const int64_t NONE = std::numeric_limits<int64_t>::max();
        DevBuf<int32_t> d_done;
        d_done.alloc(1);
        int64_t done = 0;
        // dtv[0] is the shared dt on the small path; the general path keeps it in a separate slot
        if (small_ok(true)) {
            SmallArgs<T> a;
            a.f = fd;
            a.R = R; a.Rp = Rp;
            a.v = S[cur].v.p; a.xs = S[cur].xs.p; a.xl = S[cur].xl.p;
            a.dt_arr = dtv.p;
            a.solved_step = solved.p;
            a.tol = (T)tol; a.zeta = (T)zeta; a.xl_max = T(1e4) * T(f->M);
            a.adaptive = 1;
            const size_t smem = small_smem_bytes(f->N, f->M, f->L, true, sizeof(T));
            while (max_steps < 0 || done < max_steps) {
                const int64_t k = max_steps < 0 ? (int64_t(1) << 16) : std::min<int64_t>(max_steps - done, int64_t(1) << 16);
                a.step0 = (int32_t)(step + done);
                a.nsteps = (int32_t)k;
                if (f->M > 1024) {
                    ODESAT_CUDA(cudaFuncSetAttribute(k_inter_adaptive_small<T, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmallSmem));
                    k_inter_adaptive_small<T, 1024><<<1, 1024, smem, stream>>>(a, d_done.p);
                } else {
                    ODESAT_CUDA(cudaFuncSetAttribute(k_inter_adaptive_small<T, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmallSmem));
                    k_inter_adaptive_small<T, 256><<<1, 256, smem, stream>>>(a, d_done.p);
                }
                ++launches;
                int32_t h = 0;
                ODESAT_CUDA(cudaMemcpyAsync(&h, d_done.p, 4, cudaMemcpyDeviceToHost, stream));
                ODESAT_CUDA(cudaStreamSynchronize(stream));
                ODESAT_CUDA(cudaGetLastError());
                done += h;
                if (first_key() != NONE) break;
            }
            step += done;
            return done;
        }
        // general engine: the five launches of an adaptive step on ONE replica at a time; the shared dt
        // travels through dt_shared ↔ dtv[r] around each replica's step
        ensure_alt();
        if (!H.allocated()) { H.alloc(f->N, f->M, Rp, &dev_bytes); Fb.alloc(f->N, f->M, Rp, &dev_bytes); }
        DevBuf<T> dt_shared;
        dt_shared.alloc(1);
        ODESAT_CUDA(cudaMemcpyAsync(dt_shared.p, dtv.p, sizeof(T), cudaMemcpyDeviceToDevice, stream));
        while (max_steps < 0 || done < max_steps) {
            for (int64_t r = 0; r < R; ++r) {
                ODESAT_CUDA(cudaMemcpyAsync(dtv.p + r, dt_shared.p, sizeof(T), cudaMemcpyDeviceToDevice, stream));
                GatherArgs<T> a = base_args(zeta);
                a.rep0 = r; a.rep1 = r + 1;
                a.dt_arr = dtv.p;
                a.step = (int32_t)(step + done);
                a.v = S[cur].v.p; a.xs = S[cur].xs.p; a.xl = S[cur].xl.p;
                a.ov = H.v.p; a.oxs = H.xs.p; a.oxl = H.xl.p;
                a.fv = Fb.v.p; a.fxs = Fb.xs.p; a.fxl = Fb.xl.p;
                a.yv = S[cur].v.p; a.yxs = S[cur].xs.p; a.yxl = S[cur].xl.p;
                launch_gather_v<G_ADAPT_A, 1>(a);
                a.v = H.v.p; a.xs = H.xs.p; a.xl = H.xl.p;
                a.ov = S[1 - cur].v.p; a.oxs = S[1 - cur].xs.p; a.oxl = S[1 - cur].xl.p;
                launch_gather_v<G_ADAPT_B, 1>(a);
                k_adapt_c<T><<<1, 32, 0, stream>>>(solved.p + r, unsat.p + r, err.p + r, dtv.p + r, (T)tol, 1, (int32_t)(step + done));
                ++launches;
                ODESAT_CUDA(cudaMemcpyAsync(dt_shared.p, dtv.p + r, sizeof(T), cudaMemcpyDeviceToDevice, stream));
            }
            cur = 1 - cur;
            ++done;
            v_in_range = true;
            if (first_key() != NONE) break;
        }
        ODESAT_CUDA(cudaMemcpyAsync(dtv.p, dt_shared.p, sizeof(T), cudaMemcpyDeviceToDevice, stream));
        ODESAT_CUDA(cudaStreamSynchronize(stream));
        ODESAT_CUDA(cudaGetLastError());
        step += done;
        return done;
    }

    void status(int64_t* out, int64_t* steps) override {
        if (out && R > 0) {
            std::vector<int32_t> h((size_t)R);
            ODESAT_CUDA(cudaMemcpyAsync(h.data(), solved.p, (size_t)R * 4, cudaMemcpyDeviceToHost, stream));
            ODESAT_CUDA(cudaStreamSynchronize(stream));
            for (int64_t r = 0; r < R; ++r) out[r] = h[r];
        }
        if (steps) *steps = step;
    }

    void post_key(int slot, int64_t replica_offset) override {
        k_first_key<<<1, 1024, 0, stream>>>(solved.p, R, replica_offset, d_key + 2 * slot);
        ++launches;
        ODESAT_CUDA(cudaMemcpyAsync(h_key + 2 * slot, d_key + 2 * slot, 16, cudaMemcpyDeviceToHost, stream));
        ODESAT_CUDA(cudaEventRecord(ev_key[slot], stream));
    }
    void post_key_to(int64_t replica_offset, unsigned long long* out_dev) override {
        k_first_key<<<1, 1024, 0, stream>>>(solved.p, R, replica_offset, out_dev);
        ++launches;
        ODESAT_CUDA(cudaGetLastError());
    }
    int64_t first_key() override {
        post_key(2, 0);
        int64_t k = 0;
        wait_key(2, &k, nullptr);
        return k;
    }

    ~BatchImpl() override {
        if (h_solved) cudaFreeHost(h_solved);
        if (h_bad) cudaFreeHost(h_bad);
    }
    // bad_buf[r] = replica r's thresholded state falsifies some clause (enqueue only)
    void verify_enqueue() {
        if (!bad_buf.p) bad_buf.alloc((size_t)std::max<int64_t>(R, 1), &dev_bytes);
        ODESAT_CUDA(cudaMemsetAsync(bad_buf.p, 0, bad_buf.bytes(), stream));
        bool done = false;
        if (tile && tile->has_direct() && !canon_ahead && !canon_current && f->M > 0) {
            const int64_t n = tile->verify_direct(bad_buf.p);   // on the tile layout: no export
            launches += n;
            done = n > 0;
        }
        if (!done) tile_to_canon();
        StateBuf<T>& s = canon();
        if (f->M > 0 && !done) {
            dim3 g, b;
            geom(f->M, g, b);
            k_verify<T><<<g, b, 0, stream>>>(fd, s.v.p, R, Rp, bad_buf.p);
            ++launches;
        }
        ODESAT_CUDA(cudaGetLastError());
    }
    void results_enqueue() override {
        if (R == 0) return;
        if (!h_solved) {
            ODESAT_CUDA(cudaHostAlloc((void**)&h_solved, (size_t)R * 4, cudaHostAllocDefault));
            ODESAT_CUDA(cudaHostAlloc((void**)&h_bad, (size_t)R * 4, cudaHostAllocDefault));
        }
        verify_enqueue();
        ODESAT_CUDA(cudaMemcpyAsync(h_solved, solved.p, (size_t)R * 4, cudaMemcpyDeviceToHost, stream));
        ODESAT_CUDA(cudaMemcpyAsync(h_bad, bad_buf.p, (size_t)R * 4, cudaMemcpyDeviceToHost, stream));
    }
    void results_collect(int64_t* solved_out, uint8_t* verified_out) override {
        if (R == 0) return;
        sync();
        for (int64_t r = 0; r < R; ++r) { solved_out[r] = h_solved[r]; verified_out[r] = h_bad[r] ? 0 : 1; }
    }

    void verify(uint8_t* out) override {
        if (R == 0) return;
        DevBuf<uint32_t> bad;
        bad.alloc((size_t)R);
        ODESAT_CUDA(cudaMemsetAsync(bad.p, 0, bad.bytes(), stream));
        bool done = false;
        if (tile && tile->has_direct() && !canon_ahead && !canon_current && f->M > 0) {
            const int64_t n = tile->verify_direct(bad.p);   // on the tile layout: no export
            launches += n;
            done = n > 0;
        }
        if (!done) tile_to_canon();
        StateBuf<T>& s = canon();
        if (f->M > 0 && !done) {
            dim3 g, b;
            geom(f->M, g, b);
            k_verify<T><<<g, b, 0, stream>>>(fd, s.v.p, R, Rp, bad.p);
            ++launches;
        }
        std::vector<uint32_t> h((size_t)R);
        ODESAT_CUDA(cudaMemcpyAsync(h.data(), bad.p, (size_t)R * 4, cudaMemcpyDeviceToHost, stream));
        ODESAT_CUDA(cudaStreamSynchronize(stream));
        ODESAT_CUDA(cudaGetLastError());
        for (int64_t r = 0; r < R; ++r) out[r] = h[r] ? 0 : 1;
    }

    void assignment(int64_t r, uint8_t* out) override {
        ODESAT_REQUIRE(r >= 0 && r < R, "replica index out of range");
        if (f->N == 0) return;
        if (small8.n < (size_t)f->N) small8.alloc((size_t)f->N, &dev_bytes);
        int64_t direct = 0;
        if (tile && tile->has_direct() && !canon_ahead && !canon_current) direct = tile->assignment_direct(r, small8.p);
        launches += direct;
        if (direct == 0) {
            tile_to_canon();
            StateBuf<T>& s = canon();
            k_assignment<T><<<(unsigned)((f->N + 255) / 256), 256, 0, stream>>>(s.v.p, f->N, Rp, r, small8.p);
            ++launches;
        }
        ODESAT_CUDA(cudaMemcpyAsync(out, small8.p, (size_t)f->N, cudaMemcpyDeviceToHost, stream));
        ODESAT_CUDA(cudaStreamSynchronize(stream));
        ODESAT_CUDA(cudaGetLastError());
    }

    void get_dt(double* out) override {
        std::vector<T> h((size_t)std::max<int64_t>(R, 1));
        ODESAT_CUDA(cudaMemcpyAsync(h.data(), dtv.p, (size_t)R * sizeof(T), cudaMemcpyDeviceToHost, stream));
        ODESAT_CUDA(cudaStreamSynchronize(stream));
        for (int64_t r = 0; r < R; ++r) out[r] = (double)h[r];
    }
    void set_dt(const double* in) override {
        std::vector<T> h((size_t)std::max<int64_t>(R, 1));
        for (int64_t r = 0; r < R; ++r) h[r] = (T)in[r];
        ODESAT_CUDA(cudaMemcpyAsync(dtv.p, h.data(), (size_t)R * sizeof(T), cudaMemcpyHostToDevice, stream));
        ODESAT_CUDA(cudaStreamSynchronize(stream));
    }

    // ---- single-state helpers (gather engine only) ---------------------------------------
    void derivatives(double zeta, void* dv, void* dxs, void* dxl, int* allsat) override {
        ODESAT_REQUIRE(!tile, "derivatives() needs the gather engine");
        ensure_alt();
        ODESAT_CUDA(cudaMemsetAsync(unsat.p, 0, unsat.bytes(), stream));
        GatherArgs<T> a = base_args(zeta);
        a.v = S[cur].v.p; a.xs = S[cur].xs.p; a.xl = S[cur].xl.p;
        a.ov = S[1 - cur].v.p; a.oxs = S[1 - cur].xs.p; a.oxl = S[1 - cur].xl.p;
        launch_gather<G_DERIV>(a);
        if (dv) get(dv, S[1 - cur].v.p, f->N);
        if (dxs) get(dxs, S[1 - cur].xs.p, f->M, true);
        if (dxl) get(dxl, S[1 - cur].xl.p, f->M, true);
        std::vector<uint32_t> h((size_t)std::max<int64_t>(R, 1), 0);
        ODESAT_CUDA(cudaMemcpyAsync(h.data(), unsat.p, (size_t)R * 4, cudaMemcpyDeviceToHost, stream));
        ODESAT_CUDA(cudaStreamSynchronize(stream));
        ODESAT_CUDA(cudaGetLastError());
        if (allsat) for (int64_t r = 0; r < R; ++r) allsat[r] = h[r] ? 0 : 1;
        ODESAT_CUDA(cudaMemsetAsync(unsat.p, 0, unsat.bytes(), stream));
    }

    void update_state_with(const void* dv, const void* dxs, const void* dxl, double dt) override {
        ODESAT_REQUIRE(!tile, "update_state needs the gather engine");
        ensure_alt();
        StateBuf<T>& d = S[1 - cur];
        put(dv, d.v.p, f->N);
        put(dxs, d.xs.p, f->M, true);
        put(dxl, d.xl.p, f->M, true);
        if (f->N + f->M > 0 && R > 0) {
            dim3 g, b;
            geom(f->N + f->M, g, b);
            k_update_state<T><<<g, b, 0, stream>>>(S[cur].v.p, S[cur].xs.p, S[cur].xl.p, d.v.p, d.xs.p, d.xl.p, (T)dt,
                                                    f->N, f->M, R, Rp, T(1e4) * T(f->M));
            ++launches;
        }
        ODESAT_CUDA(cudaStreamSynchronize(stream));
        ODESAT_CUDA(cudaGetLastError());
    }

    double max_error_vs(const void* bv, const void* bxs, const void* bxl) override {
        ODESAT_REQUIRE(!tile && R == 1, "max_error is a single-state call");
        ensure_alt();
        StateBuf<T>& d = S[1 - cur];
        put(bv, d.v.p, f->N);
        put(bxs, d.xs.p, f->M, true);
        put(bxl, d.xl.p, f->M, true);
        ODESAT_CUDA(cudaMemsetAsync(err.p, 0, err.bytes(), stream));
        if (f->N + f->M > 0) {
            dim3 g, b;
            geom(f->N + f->M, g, b);
            k_max_error<T><<<g, b, 0, stream>>>(S[cur].v.p, S[cur].xs.p, S[cur].xl.p, d.v.p, d.xs.p, d.xl.p, f->N, f->M,
                                                 R, Rp, err.p);
            ++launches;
        }
        if (dscratch.n < 1) dscratch.alloc(1, &dev_bytes);
        k_err_decode<T><<<1, 32, 0, stream>>>(err.p, dscratch.p, 1);
        ++launches;
        double h = 0;
        ODESAT_CUDA(cudaMemcpyAsync(&h, dscratch.p, 8, cudaMemcpyDeviceToHost, stream));
        ODESAT_CUDA(cudaMemsetAsync(err.p, 0, err.bytes(), stream));
        ODESAT_CUDA(cudaStreamSynchronize(stream));
        ODESAT_CUDA(cudaGetLastError());
        return h;
    }
};

}  // namespace odesat

struct odesat_batch {
    std::unique_ptr<odesat::BatchBase> impl;
};
