// odesat_b200.cu — implementation of the C ABI declared in include/odesat_b200.h.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo -shared
// There is no CPU fallback in this file: every compute entry point needs a CUDA device.
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "batch.cuh"

using namespace odesat;

namespace {

void require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        throw Error(ODESAT_ECUDA, "no CUDA device available (odesat_b200 has no CPU fallback)");
    }
}

BatchBase* make_batch(const odesat_formula* f, int64_t R, int precision, int engine, int schedule) {
    ODESAT_REQUIRE(f != nullptr, "formula is NULL");
    ODESAT_REQUIRE(R >= 0, "negative replica count");
    ODESAT_REQUIRE(precision == ODESAT_F64 || precision == ODESAT_F32, "unknown precision");
    ODESAT_REQUIRE(engine >= ODESAT_ENGINE_AUTO && engine <= ODESAT_ENGINE_TILE, "unknown engine");
    ODESAT_REQUIRE(schedule == ODESAT_SCHED_EXACT || schedule == ODESAT_SCHED_BALANCED, "unknown schedule");
    require_device();
    int dev = 0;
    ODESAT_CUDA(cudaGetDevice(&dev));
    ODESAT_REQUIRE(dev == f->device, "formula was created on a different CUDA device");
    if (precision == ODESAT_F32) return new BatchImpl<float>(f, R, engine, schedule);
    return new BatchImpl<double>(f, R, engine, schedule);
}

// Batch whose device buffers live on the formula handle and are reused by the next
// odesat_simulate* call of the same shape.
BatchBase* cached_batch(const odesat_formula* f, int64_t R, int precision, int engine, int schedule) {
    ODESAT_REQUIRE(f != nullptr, "formula is NULL");
    if (f->batch_cache && f->cache_R == R && f->cache_precision == precision && f->cache_engine == engine &&
        f->cache_schedule == schedule) {
        BatchBase* b = static_cast<BatchBase*>(f->batch_cache.get());
        b->reset();
        return b;
    }
    f->batch_cache.reset();   // free the old buffers before allocating the new ones
    BatchBase* b = make_batch(f, R, precision, engine, schedule);
    f->batch_cache = std::shared_ptr<void>(b, [](void* p) { delete static_cast<BatchBase*>(p); });
    f->cache_R = R; f->cache_precision = precision; f->cache_engine = engine; f->cache_schedule = schedule;
    return b;
}

// Host buffers of type TH feeding a device batch of precision `prec`.
template <typename TH> struct HostIO {
    int prec;
    std::vector<float> f32[3];
    std::vector<double> f64[3];
    explicit HostIO(int p) : prec(p) {}
    bool same() const { return (prec == ODESAT_F32) == (sizeof(TH) == 4); }
    const void* in(int slot, const TH* p, size_t n) {
        if (!p) return nullptr;
        if (same()) return p;
        if (prec == ODESAT_F32) { f32[slot].assign(p, p + n); return f32[slot].data(); }
        f64[slot].assign(p, p + n);
        return f64[slot].data();
    }
    void* out_buf(int slot, TH* p, size_t n) {
        if (!p) return nullptr;
        if (same()) return p;
        if (prec == ODESAT_F32) { f32[slot].resize(n); return f32[slot].data(); }
        f64[slot].resize(n);
        return f64[slot].data();
    }
    void out_commit(int slot, TH* p, size_t n) {
        if (!p || same()) return;
        if (prec == ODESAT_F32) for (size_t i = 0; i < n; ++i) p[i] = (TH)f32[slot][i];
        else for (size_t i = 0; i < n; ++i) p[i] = (TH)f64[slot][i];
    }
};

template <typename TH>
void upload_host(BatchBase& b, const TH* v, const TH* xs, const TH* xl, bool reset) {
    HostIO<TH> io(b.precision);
    const size_t nv = (size_t)(b.R * b.f->N), nm = (size_t)(b.R * b.f->M);
    b.upload(io.in(0, v, nv), io.in(1, xs, nm), io.in(2, xl, nm), reset);
}
template <typename TH> void download_host(BatchBase& b, TH* v, TH* xs, TH* xl) {
    HostIO<TH> io(b.precision);
    const size_t nv = (size_t)(b.R * b.f->N), nm = (size_t)(b.R * b.f->M);
    void* pv = io.out_buf(0, v, nv);
    void* ps = io.out_buf(1, xs, nm);
    void* pl = io.out_buf(2, xl, nm);
    b.download(pv, ps, pl);
    io.out_commit(0, v, nv);
    io.out_commit(1, xs, nm);
    io.out_commit(2, xl, nm);
}

struct Resolved {
    bool fixed;
    double dt, tol, zeta;
    int64_t steps;
    int chunk;
};
Resolved resolve(const odesat_formula* f, const odesat_params* p) {
    ODESAT_REQUIRE(p != nullptr, "params is NULL");
    Resolved r;
    r.fixed = !std::isnan(p->step_size);                       // system.rs:190
    r.dt = r.fixed ? p->step_size : 0.01;                      // system.rs:205
    r.tol = std::isnan(p->tolerance) ? 1e-3 : p->tolerance;    // system.rs:174
    r.zeta = std::isnan(p->learning_rate) ? f->default_zeta() : p->learning_rate;   // system.rs:164-173
    r.steps = p->steps;
    r.chunk = p->chunk > 0 ? p->chunk : 32;
    return r;
}

// The step loop of simulate / batch / simulate_inter with chunked early-exit polling.
// BATCH stops when every replica has flagged; INTER when any has.  Returns the INTER key
// (INT64_MAX when no replica flagged).
int64_t drive(BatchBase& b, const Resolved& r, int mode, std::vector<int64_t>& solved) {
    const int64_t NONE = std::numeric_limits<int64_t>::max();
    solved.assign((size_t)b.R, -1);
    int64_t key = NONE;
    if (b.R == 0) return key;
    int64_t done_steps = 0;
    while (r.steps < 0 || done_steps < r.steps) {
        const int64_t n = r.steps < 0 ? r.chunk : std::min<int64_t>(r.chunk, r.steps - done_steps);
        if (r.fixed) b.run_fixed(r.dt, r.zeta, n, /*freeze=*/1, nullptr);
        else b.run_adaptive(r.tol, r.zeta, n, nullptr);
        done_steps += n;
        if (mode == ODESAT_MODE_INTER) {
            key = b.first_key();
            if (key != NONE) break;
        } else {
            b.status(solved.data(), nullptr);
            bool all = true;
            for (int64_t x : solved) all = all && x >= 0;
            if (all) break;
        }
    }
    b.status(solved.data(), nullptr);
    if (mode == ODESAT_MODE_INTER && key == NONE) key = b.first_key();
    return key;
}

template <typename TH>
void simulate_impl(const odesat_formula* f, TH* v, TH* xs, TH* xl, const odesat_params* p, uint8_t* assignment,
                   int64_t* steps_taken, int* allsat, double* final_dt) {
    ODESAT_REQUIRE(f && v && xs && xl, "NULL state or formula");
    Resolved r = resolve(f, p);
    const int eng = (!r.fixed && p->engine == ODESAT_ENGINE_AUTO) ? ODESAT_ENGINE_GATHER : p->engine;
    std::unique_ptr<BatchBase> b(make_batch(f, 1, p->precision, eng, p->schedule));
    if (p->chunk <= 0) r.chunk = b->preferred_chunk();
    upload_host<TH>(*b, v, xs, xl, true);
    std::vector<int64_t> solved;
    drive(*b, r, ODESAT_MODE_BATCH, solved);
    download_host<TH>(*b, v, xs, xl);
    if (assignment) b->assignment(0, assignment);                                // system.rs:238
    if (steps_taken) *steps_taken = solved[0] >= 0 ? solved[0] + 1 : (r.steps < 0 ? b->step : r.steps);
    if (allsat) *allsat = solved[0] >= 0 ? 1 : 0;
    if (final_dt) { double d = r.dt; if (!r.fixed) b->get_dt(&d); *final_dt = d; }
}

template <typename TH>
void simulate_batch_impl(const odesat_formula* f, int64_t R, TH* v, TH* xs, TH* xl, uint64_t seed,
                         int64_t replica_offset, const odesat_params* p, int mode, int write_back,
                         int64_t* solved_step, uint8_t* verified, int64_t* winner, uint8_t* assignment,
                         int64_t* steps_run) {
    ODESAT_REQUIRE(f != nullptr, "formula is NULL");
    ODESAT_REQUIRE(mode == ODESAT_MODE_BATCH || mode == ODESAT_MODE_INTER, "unknown mode");
    Resolved r = resolve(f, p);
    if (mode == ODESAT_MODE_BATCH) ODESAT_REQUIRE(r.steps >= 0, "batch needs a step count (main.rs:96-97)");
    // the tile engine integrates fixed steps only: adaptive runs resolve AUTO to the gather engine
    const int eng = (!r.fixed && p->engine == ODESAT_ENGINE_AUTO) ? ODESAT_ENGINE_GATHER : p->engine;
    BatchBase* b = cached_batch(f, R, p->precision, eng, p->schedule);
    if (p->chunk <= 0) r.chunk = b->preferred_chunk();
    const int64_t NONE = std::numeric_limits<int64_t>::max();
    // main.rs:283-289: whatever the caller does not supply is generated on the device
    if (!(v && xs && xl)) b->init(seed, replica_offset, !v, !xs, !xl, /*finalize=*/!(v || xs || xl));
    if (v || xs || xl) upload_host<TH>(*b, v, xs, xl, (v && xs && xl));
    std::vector<int64_t> solved;
    int64_t key;
    if (mode == ODESAT_MODE_INTER && !r.fixed) {
        // adaptive inter: the replicas share ONE dt and step one after the other (system.rs:312-349, quirk Q7)
        b->run_inter_adaptive(r.tol, r.zeta, r.steps);
        solved.assign((size_t)R, -1);
        b->status(solved.data(), nullptr);
        key = b->first_key();
    } else {
        key = drive(*b, r, mode, solved);
    }
    std::vector<uint8_t> ver((size_t)std::max<int64_t>(R, 1), 0);
    b->verify(ver.data());                                                       // cnf.rs:246-264
    int64_t win = -1, src = 0;
    int64_t run = b->step;
    if (mode == ODESAT_MODE_BATCH) {
        for (int64_t q = 0; q < R; ++q) if (ver[q]) { win = q; break; }          // main.rs:305-307
        src = win >= 0 ? win : R - 1;
    } else {
        if (r.steps == 0) win = R > 0 ? 0 : -1;                                  // system.rs:274, 353 (Q8)
        else if (key != NONE) { win = key & 0xFFFFFFFFll; run = (key >> 32) + 1; }
        src = win >= 0 ? win : 0;                                                // system.rs:357
    }
    if (solved_step) for (int64_t q = 0; q < R; ++q) solved_step[q] = solved[q];
    if (verified) for (int64_t q = 0; q < R; ++q) verified[q] = ver[q];
    if (winner) *winner = win;
    if (assignment && R > 0) b->assignment(src, assignment);
    if (steps_run) *steps_run = run;
    if (write_back) download_host<TH>(*b, v, xs, xl);
}

template <typename TH> std::unique_ptr<BatchBase> single(const odesat_formula* f, const TH* v, const TH* xs, const TH* xl) {
    ODESAT_REQUIRE(f && v && xs && xl, "NULL state or formula");
    std::unique_ptr<BatchBase> b(make_batch(f, 1, sizeof(TH) == 4 ? ODESAT_F32 : ODESAT_F64, ODESAT_ENGINE_GATHER,
                                            ODESAT_SCHED_EXACT));
    b->upload(v, xs, xl, true);
    return b;
}

template <typename TH>
void derivatives_impl(const odesat_formula* f, const TH* v, const TH* xs, const TH* xl, double zeta, TH* dv, TH* dxs,
                      TH* dxl, int* allsat) {
    auto b = single<TH>(f, v, xs, xl);
    int a = 0;
    b->derivatives(zeta, dv, dxs, dxl, &a);
    if (allsat) *allsat = a;
}
template <typename TH>
void update_impl(const odesat_formula* f, TH* v, TH* xs, TH* xl, const TH* dv, const TH* dxs, const TH* dxl, double dt) {
    ODESAT_REQUIRE(dv && dxs && dxl, "NULL derivative");
    auto b = single<TH>(f, v, xs, xl);
    b->update_state_with(dv, dxs, dxl, dt);
    b->download(v, xs, xl);
}
template <typename TH>
void maxerr_impl(const odesat_formula* f, const TH* av, const TH* axs, const TH* axl, const TH* bv, const TH* bxs,
                 const TH* bxl, double* err) {
    ODESAT_REQUIRE(bv && bxs && bxl && err, "NULL argument");
    auto b = single<TH>(f, av, axs, axl);
    *err = b->max_error_vs(bv, bxs, bxl);
}
template <typename TH>
void step_fixed_impl(const odesat_formula* f, TH* v, TH* xs, TH* xl, double dt, double zeta, int* allsat) {
    auto b = single<TH>(f, v, xs, xl);
    b->run_fixed(dt, zeta, 1, /*freeze=*/0, nullptr);
    int64_t s = -1;
    b->status(&s, nullptr);
    b->download(v, xs, xl);
    if (allsat) *allsat = s == 0 ? 1 : 0;
}
template <typename TH>
void step_adaptive_impl(const odesat_formula* f, TH* v, TH* xs, TH* xl, double tol, double* dt, double zeta, int* allsat) {
    ODESAT_REQUIRE(dt != nullptr, "dt is NULL");
    auto b = single<TH>(f, v, xs, xl);
    b->set_dt(dt);
    b->run_adaptive(tol, zeta, 1, nullptr);
    int64_t s = -1;
    b->status(&s, nullptr);
    b->download(v, xs, xl);
    b->get_dt(dt);
    if (allsat) *allsat = s == 0 ? 1 : 0;
}
template <typename TH> void xs0_impl(const odesat_formula* f, TH* xs0) {
    ODESAT_REQUIRE(f && xs0, "NULL argument");
    for (int64_t m = 0; m < f->M; ++m) xs0[m] = (TH)f->h_xs0[m];   // system.rs:362-372 (host-side, O(L) at create)
}

}  // namespace

extern "C" {

const char* odesat_last_error(void) { return last_error_ref().c_str(); }
int odesat_abi_version(void) { return ODESAT_B200_ABI_VERSION; }
int odesat_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int odesat_formula_create(int64_t varnum, int64_t n_clauses, const int64_t* clause_off, const int32_t* lits,
                          odesat_formula** out) {
    return guarded([&] {
        ODESAT_REQUIRE(out != nullptr, "out is NULL");
        *out = nullptr;
        std::unique_ptr<odesat_formula> f(new odesat_formula);
        f->build(varnum, n_clauses, clause_off, lits);
        require_device();
        f->upload();
        *out = f.release();
    });
}
void odesat_formula_destroy(odesat_formula* f) { delete f; }
int odesat_formula_info(const odesat_formula* f, int64_t* varnum, int64_t* n_clauses, int64_t* n_literals,
                        int32_t* uniform_k) {
    return guarded([&] {
        ODESAT_REQUIRE(f != nullptr, "formula is NULL");
        if (varnum) *varnum = f->N;
        if (n_clauses) *n_clauses = f->M;
        if (n_literals) *n_literals = f->L;
        if (uniform_k) *uniform_k = f->K;
    });
}
int odesat_formula_default_zeta(const odesat_formula* f, double* zeta) {
    return guarded([&] {
        ODESAT_REQUIRE(f && zeta, "NULL argument");
        *zeta = f->default_zeta();
    });
}

int odesat_init_short_term_memory(const odesat_formula* f, double* xs0) { return guarded([&] { xs0_impl<double>(f, xs0); }); }
int odesat_init_short_term_memory_f32(const odesat_formula* f, float* xs0) { return guarded([&] { xs0_impl<float>(f, xs0); }); }

int odesat_compute_derivatives(const odesat_formula* f, const double* v, const double* xs, const double* xl, double zeta,
                               double* dv, double* dxs, double* dxl, int* allsat) {
    return guarded([&] { derivatives_impl<double>(f, v, xs, xl, zeta, dv, dxs, dxl, allsat); });
}
int odesat_compute_derivatives_f32(const odesat_formula* f, const float* v, const float* xs, const float* xl, double zeta,
                                   float* dv, float* dxs, float* dxl, int* allsat) {
    return guarded([&] { derivatives_impl<float>(f, v, xs, xl, zeta, dv, dxs, dxl, allsat); });
}
int odesat_update_state(const odesat_formula* f, double* v, double* xs, double* xl, const double* dv, const double* dxs,
                        const double* dxl, double dt) {
    return guarded([&] { update_impl<double>(f, v, xs, xl, dv, dxs, dxl, dt); });
}
int odesat_update_state_f32(const odesat_formula* f, float* v, float* xs, float* xl, const float* dv, const float* dxs,
                            const float* dxl, double dt) {
    return guarded([&] { update_impl<float>(f, v, xs, xl, dv, dxs, dxl, dt); });
}
int odesat_max_error(const odesat_formula* f, const double* av, const double* axs, const double* axl, const double* bv,
                     const double* bxs, const double* bxl, double* err) {
    return guarded([&] { maxerr_impl<double>(f, av, axs, axl, bv, bxs, bxl, err); });
}
int odesat_max_error_f32(const odesat_formula* f, const float* av, const float* axs, const float* axl, const float* bv,
                         const float* bxs, const float* bxl, double* err) {
    return guarded([&] { maxerr_impl<float>(f, av, axs, axl, bv, bxs, bxl, err); });
}
int odesat_euler_step_fixed(const odesat_formula* f, double* v, double* xs, double* xl, double dt, double zeta, int* allsat) {
    return guarded([&] { step_fixed_impl<double>(f, v, xs, xl, dt, zeta, allsat); });
}
int odesat_euler_step_fixed_f32(const odesat_formula* f, float* v, float* xs, float* xl, double dt, double zeta, int* allsat) {
    return guarded([&] { step_fixed_impl<float>(f, v, xs, xl, dt, zeta, allsat); });
}
int odesat_euler_step(const odesat_formula* f, double* v, double* xs, double* xl, double tolerance, double* dt, double zeta,
                      int* allsat) {
    return guarded([&] { step_adaptive_impl<double>(f, v, xs, xl, tolerance, dt, zeta, allsat); });
}
int odesat_euler_step_f32(const odesat_formula* f, float* v, float* xs, float* xl, double tolerance, double* dt, double zeta,
                          int* allsat) {
    return guarded([&] { step_adaptive_impl<float>(f, v, xs, xl, tolerance, dt, zeta, allsat); });
}
int odesat_simulate(const odesat_formula* f, double* v, double* xs, double* xl, const odesat_params* params,
                    uint8_t* assignment, int64_t* steps_taken, int* allsat, double* final_dt) {
    return guarded([&] { simulate_impl<double>(f, v, xs, xl, params, assignment, steps_taken, allsat, final_dt); });
}
int odesat_simulate_f32(const odesat_formula* f, float* v, float* xs, float* xl, const odesat_params* params,
                        uint8_t* assignment, int64_t* steps_taken, int* allsat, double* final_dt) {
    return guarded([&] { simulate_impl<float>(f, v, xs, xl, params, assignment, steps_taken, allsat, final_dt); });
}
int odesat_simulate_batch(const odesat_formula* f, int64_t R, double* v, double* xs, double* xl, uint64_t seed,
                          int64_t replica_offset, const odesat_params* params, int32_t mode, int32_t write_back,
                          int64_t* solved_step, uint8_t* verified, int64_t* winner, uint8_t* assignment, int64_t* steps_run) {
    return guarded([&] {
        simulate_batch_impl<double>(f, R, v, xs, xl, seed, replica_offset, params, mode, write_back, solved_step, verified,
                                    winner, assignment, steps_run);
    });
}
int odesat_simulate_batch_f32(const odesat_formula* f, int64_t R, float* v, float* xs, float* xl, uint64_t seed,
                              int64_t replica_offset, const odesat_params* params, int32_t mode, int32_t write_back,
                              int64_t* solved_step, uint8_t* verified, int64_t* winner, uint8_t* assignment,
                              int64_t* steps_run) {
    return guarded([&] {
        simulate_batch_impl<float>(f, R, v, xs, xl, seed, replica_offset, params, mode, write_back, solved_step, verified,
                                   winner, assignment, steps_run);
    });
}
int odesat_simulate_inter(const odesat_formula* f, int64_t R, double* v, double* xs, double* xl, const odesat_params* params,
                          uint8_t* assignment, int64_t* winner, int64_t* steps_taken) {
    return guarded([&] {
        ODESAT_REQUIRE(v && xs && xl, "simulate_inter takes caller-supplied states");
        simulate_batch_impl<double>(f, R, v, xs, xl, 0, 0, params, ODESAT_MODE_INTER, 1, nullptr, nullptr, winner, assignment,
                                    steps_taken);
    });
}

int odesat_tile_schedule_stats(int64_t varnum, int64_t n_clauses, const int64_t* clause_off, const int32_t* lits,
                               int32_t schedule, int32_t threads, int32_t depth, int32_t* perm, int64_t perm_capacity,
                               uint32_t* items, int64_t items_capacity, int64_t* out) {
    return guarded([&] {
        ODESAT_REQUIRE(out != nullptr, "out is NULL");
        ODESAT_REQUIRE(schedule == ODESAT_SCHED_EXACT || schedule == ODESAT_SCHED_BALANCED, "unknown schedule");
        ODESAT_REQUIRE(threads >= 32 && threads <= 1024 && threads % 32 == 0 && depth >= 1 && depth <= 16, "bad threads/depth");
        odesat_formula f;
        f.build(varnum, n_clauses, clause_off, lits);
        ODESAT_REQUIRE(f.K == 3 && f.distinct_vars, "tile schedules need uniform 3-literal clauses with distinct variables");
        // the same level construction the tile engine uses for a CTA of `threads` threads
        const int target = threads >= 512 ? threads : 1024;
        auto lv = schedule == ODESAT_SCHED_EXACT ? build_tile_levels(f, schedule, threads) : build_tile_levels(f, schedule, target, target / 2);
        auto s = build_tile_schedule(f, *lv, schedule, threads, depth, /*upload=*/false);
        out[0] = s->nlev;
        out[1] = s->n_items;
        out[2] = s->Mpad;
        out[3] = (int64_t)(s->conflict_wavefronts * 1000.0 + 0.5);
        if (perm) {
            ODESAT_REQUIRE(perm_capacity >= s->Mpad, "perm buffer too small");
            for (int64_t i = 0; i < s->Mpad; ++i) perm[i] = s->perm[i];
        }
        if (items) {
            ODESAT_REQUIRE(items_capacity >= s->n_items, "items buffer too small");
            for (int64_t i = 0; i < s->n_items; ++i) items[i] = s->items[i];
        }
    });
}

int odesat_batch_create(const odesat_formula* f, int64_t R, int32_t precision, int32_t engine, int32_t schedule,
                        odesat_batch** out) {
    return guarded([&] {
        ODESAT_REQUIRE(out != nullptr, "out is NULL");
        *out = nullptr;
        std::unique_ptr<odesat_batch> b(new odesat_batch);
        b->impl.reset(make_batch(f, R, precision, engine, schedule));
        *out = b.release();
    });
}
void odesat_batch_destroy(odesat_batch* b) { delete b; }
int odesat_batch_info(const odesat_batch* b, int32_t* engine, int64_t* kernel_launches, int64_t* device_bytes) {
    return guarded([&] {
        ODESAT_REQUIRE(b != nullptr, "batch is NULL");
        if (engine) *engine = b->impl->engine;
        if (kernel_launches) *kernel_launches = b->impl->launches;
        if (device_bytes) *device_bytes = b->impl->dev_bytes;
    });
}
int odesat_batch_init(odesat_batch* b, uint64_t seed, int64_t replica_offset) {
    return guarded([&] {
        ODESAT_REQUIRE(b != nullptr, "batch is NULL");
        b->impl->reset();                                    // flags / step / dt of a fresh batch
        b->impl->init(seed, replica_offset, true, true, true);
    });
}
int odesat_batch_upload(odesat_batch* b, const void* v, const void* xs, const void* xl) {
    return guarded([&] {
        ODESAT_REQUIRE(b && v && xs && xl, "NULL argument");
        b->impl->upload(v, xs, xl, true);
    });
}
int odesat_batch_download(odesat_batch* b, void* v, void* xs, void* xl) {
    return guarded([&] {
        ODESAT_REQUIRE(b != nullptr, "batch is NULL");
        b->impl->download(v, xs, xl);
    });
}
int odesat_batch_run_fixed(odesat_batch* b, double dt, double zeta, int64_t n, int32_t freeze, float* device_ms) {
    return guarded([&] {
        ODESAT_REQUIRE(b != nullptr, "batch is NULL");
        b->impl->run_fixed(dt, zeta, n, freeze, device_ms);
    });
}
int odesat_batch_run_adaptive(odesat_batch* b, double tolerance, double zeta, int64_t n, float* device_ms) {
    return guarded([&] {
        ODESAT_REQUIRE(b != nullptr, "batch is NULL");
        b->impl->run_adaptive(tolerance, zeta, n, device_ms);
    });
}
int odesat_batch_status(odesat_batch* b, int64_t* solved_step, int64_t* steps_done) {
    return guarded([&] {
        ODESAT_REQUIRE(b != nullptr, "batch is NULL");
        b->impl->status(solved_step, steps_done);
    });
}
int odesat_batch_first_solved(odesat_batch* b, int64_t* key) {
    return guarded([&] {
        ODESAT_REQUIRE(b && key, "NULL argument");
        *key = b->impl->first_key();
    });
}
int odesat_batch_verify(odesat_batch* b, uint8_t* verified) {
    return guarded([&] {
        ODESAT_REQUIRE(b && verified, "NULL argument");
        b->impl->verify(verified);
    });
}
int odesat_batch_assignment(odesat_batch* b, int64_t replica, uint8_t* assignment) {
    return guarded([&] {
        ODESAT_REQUIRE(b && assignment, "NULL argument");
        b->impl->assignment(replica, assignment);
    });
}
int odesat_batch_dt(odesat_batch* b, double* dt) {
    return guarded([&] {
        ODESAT_REQUIRE(b && dt, "NULL argument");
        b->impl->get_dt(dt);
    });
}

}  // extern "C"
