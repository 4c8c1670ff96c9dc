// odesat_b200.cu — implementation of the C ABI declared in include/odesat_b200.h.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo -shared
// There is no CPU fallback in this file: every compute entry point needs a CUDA device.
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <functional>
#include <limits>
#include <string>
#include <vector>

#include "batch.cuh"
#include "stoch.cuh"

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

// RAII: makes `dev` current, restores the previous device on scope exit
struct DeviceGuard {
    int prev = 0;
    explicit DeviceGuard(int dev) {
        ODESAT_CUDA(cudaGetDevice(&prev));
        if (dev != prev) ODESAT_CUDA(cudaSetDevice(dev));
    }
    ~DeviceGuard() { cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// A batch on the device that is current (must be the formula's own device).
BatchBase* make_batch(const odesat_formula* f, int64_t R, int precision, int engine, int schedule) {
    ODESAT_REQUIRE(f != nullptr, "formula is NULL");
    ODESAT_REQUIRE(R >= 0, "negative replica count");
    ODESAT_REQUIRE(precision == ODESAT_F64 || precision == ODESAT_F32, "unknown precision");
    ODESAT_REQUIRE((engine >= ODESAT_ENGINE_AUTO && engine <= ODESAT_ENGINE_SLAB) || engine == ENGINE_AUTO_ADAPTIVE, "unknown engine");
    ODESAT_REQUIRE(schedule == ODESAT_SCHED_EXACT || schedule == ODESAT_SCHED_BALANCED, "unknown schedule");
    require_device();
    int dev = 0;
    ODESAT_CUDA(cudaGetDevice(&dev));
    ODESAT_REQUIRE(dev == f->device, "formula was created on a different CUDA device");
    if (precision == ODESAT_F32) return new BatchImpl<float>(f, R, engine, schedule);
    return new BatchImpl<double>(f, R, engine, schedule);
}

// One shard of a replica-batch call: replicas [off, off + R) of the call on one device, with its own stream.
struct Shard {
    BatchBase* b = nullptr;
    int dev = 0;
    int64_t off = 0, R = 0;
};

// Contiguous replica ranges: device g of G owns [g·R/G, (g+1)·R/G); a device's range is cut the
// same way into `sub` sub-batches.  Devices: the formula's own first, then the others in ascending order.
std::vector<Shard> plan_shards(const odesat_formula* f, int64_t R, int n_gpus, int sub) {
    std::vector<int> devs{f->device};
    for (int d = 0; (int)devs.size() < n_gpus; ++d) if (d != f->device) devs.push_back(d);
    std::vector<Shard> out;
    for (int g = 0; g < n_gpus; ++g) {
        const int64_t lo = R * g / n_gpus, hi = R * (g + 1) / n_gpus;
        for (int k = 0; k < sub; ++k) {
            Shard s;
            s.dev = devs[g];
            s.off = lo + (hi - lo) * k / sub;
            s.R = lo + (hi - lo) * (k + 1) / sub - s.off;
            if (s.R > 0 || (g == 0 && k == 0)) out.push_back(s);   // an empty call keeps one (empty) shard
        }
    }
    return out;
}

// Batches whose device buffers live on the formula handle and are reused by the next odesat_simulate* call of
// the same shape.
void cached_shards(const odesat_formula* f, std::vector<Shard>& sh, int precision, int engine, int schedule) {
    std::string key = std::to_string(precision) + "/" + std::to_string(engine) + "/" + std::to_string(schedule);
    for (const Shard& s : sh) key += "|" + std::to_string(s.dev) + ":" + std::to_string(s.off) + "+" + std::to_string(s.R);
    if (f->cache_key != key || f->batch_cache.size() != sh.size()) {
        f->batch_cache.clear();   // free the old buffers before allocating the new ones
        f->cache_key.clear();
        for (Shard& s : sh) {
            DeviceGuard g(s.dev);
            BatchBase* b = make_batch(f->on_device(s.dev), s.R, precision, engine, schedule);
            f->batch_cache.emplace_back(std::shared_ptr<void>(b, [](void* p) { delete static_cast<BatchBase*>(p); }));
        }
        f->cache_key = key;
    }
    for (size_t i = 0; i < sh.size(); ++i) {
        sh[i].b = static_cast<BatchBase*>(f->batch_cache[i].get());
        DeviceGuard g(sh[i].dev);
        sh[i].b->reset();
    }
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

// Enqueues the upload; returns without waiting when the host type is the device precision (the caller's buffers
// stay valid until the call's final sync), after a sync when a converted temporary was used.
template <typename TH>
void upload_host(BatchBase& b, const TH* v, const TH* xs, const TH* xl, bool reset) {
    HostIO<TH> io(b.precision);
    const size_t nv = (size_t)(b.R * b.f->N), nm = (size_t)(b.R * b.f->M);
    b.upload(io.in(0, v, nv), io.in(1, xs, nm), io.in(2, xl, nm), reset);
    if (!io.same()) b.sync();
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

constexpr int64_t NO_KEY = std::numeric_limits<int64_t>::max();

struct DriveResult {
    int64_t key = NO_KEY;   // INTER: min over all replicas of (first flagged step << 32 | replica index in the call)
    int64_t run = 0;        // Euler steps of the loop up to and including the chunk that ended it
};

// The step loop of simulate / batch / simulate_inter over the shards of one call.
//
// Chunks of r.chunk steps are ENQUEUED on every shard's stream, each followed by the reduction of the shard's flags
// into a pinned host slot (BatchBase::post_key); the host reads the slots of chunk c only after it has issued chunk
// c + 1, so no stream ever drains while the host decides.  BATCH ends when no replica is left unflagged, INTER when
// some replica has flagged.  The speculatively issued chunk is harmless: flagged replicas are frozen (freeze = 1),
// and INTER kernels read the previous chunk's key on the device and do nothing once it is set.
//
// lockstep (INTER with fixed steps whose states are written back): the reference leaves EVERY replica after exactly
// the winning step (system.rs:279-293), so a chunk is polled synchronously and, when it contains the winning step
// s*, replayed from a snapshot of its start for exactly s* − start + 1 steps without freezing.
//
// `prepare(k)` enqueues shard k's inputs; it is called right before the shard's first chunk, so the host→device
// copy of shard k + 1 overlaps the integration of shard k.
// `after_final(k)` (may be empty) is called right after shard k's FINAL chunk has been enqueued (the step budget is then
// exhausted whatever the flags say): the caller enqueues the shard's result kernels there, so that they run while the
// other shards still integrate instead of after every shard has drained.
template <typename Prepare>
DriveResult drive(std::vector<Shard>& sh, const Resolved& r, int mode, bool lockstep, Prepare&& prepare,
                  const std::function<void(size_t)>& after_final = nullptr) {
    DriveResult out;
    const bool inter = mode == ODESAT_MODE_INTER;
    bool prepared = false;
    int64_t total = 0;
    for (const Shard& s : sh) total += s.R;
    auto poll = [&](int slot, int64_t end_steps) {
        int64_t key = NO_KEY, unflagged = 0;
        for (Shard& s : sh) {
            int64_t k = NO_KEY, u = 0;
            DeviceGuard g(s.dev);
            s.b->wait_key(slot, &k, &u);
            key = std::min(key, k);
            unflagged += u;
        }
        if (inter ? key != NO_KEY : unflagged == 0) {
            out.key = key;
            out.run = end_steps;
            return true;
        }
        out.key = key;
        return false;
    };
    int64_t issued = 0, last_chunk = -1;
    bool stopped = false;
    if (total > 0) {
        for (int64_t c = 0; r.steps < 0 || issued < r.steps; ++c) {
            last_chunk = c;
            const int64_t n = r.steps < 0 ? r.chunk : std::min<int64_t>(r.chunk, r.steps - issued);
            const int slot = lockstep ? 0 : (int)(c & 1);
            for (size_t k = 0; k < sh.size(); ++k) {
                Shard& s = sh[k];
                DeviceGuard g(s.dev);
                if (c == 0) prepare(k);
                if (lockstep && n > 1) s.b->snapshot();
                const unsigned long long* stop = (inter && !lockstep && c > 0) ? s.b->key_dev((int)((c - 1) & 1)) : nullptr;
                if (r.fixed) s.b->run_fixed_async(r.dt, r.zeta, n, /*freeze=*/1, stop);
                else s.b->run_adaptive_async(r.tol, r.zeta, n);
                s.b->post_key(slot, s.off);
                if (after_final && !lockstep && r.steps >= 0 && issued + n >= r.steps) after_final(k);
            }
            prepared = true;
            issued += n;
            if (lockstep) {
                if (poll(0, issued)) {
                    const int64_t s_star = out.key >> 32;
                    if (n > 1) {   // replay the chunk up to and including the winning step, nobody frozen
                        const int64_t need = s_star - (issued - n) + 1;
                        for (Shard& s : sh) {
                            DeviceGuard g(s.dev);
                            s.b->restore();
                            s.b->run_fixed_async(r.dt, r.zeta, need, /*freeze=*/0, nullptr);
                            s.b->post_key(0, s.off);
                        }
                        poll(0, issued);
                    }
                    out.run = s_star + 1;
                    stopped = true;
                    break;
                }
            } else if (c >= 1 && poll((int)((c - 1) & 1), issued - n)) {
                stopped = true;
                break;
            }
        }
        if (!stopped && !lockstep && last_chunk >= 0) stopped = poll((int)(last_chunk & 1), issued);
        if (!stopped) out.run = issued;
        if (inter && out.key != NO_KEY) out.run = (out.key >> 32) + 1;
    }
    if (!prepared) for (size_t k = 0; k < sh.size(); ++k) { DeviceGuard g(sh[k].dev); prepare(k); }
    for (Shard& s : sh) { DeviceGuard g(s.dev); s.b->sync(); }
    return out;
}

template <typename TH>
void simulate_impl(const odesat_formula* f, TH* v, TH* xs, TH* xl, const odesat_params* p, uint8_t* assignment,
                   int64_t* steps_taken, int* allsat, double* final_dt) {
    ODESAT_REQUIRE(f && v && xs && xl, "NULL state or formula");
    Resolved r = resolve(f, p);
    const int eng = (!r.fixed && p->engine == ODESAT_ENGINE_AUTO) ? ODESAT_ENGINE_GATHER : p->engine;
    DeviceGuard guard(f->device);
    std::unique_ptr<BatchBase> b(make_batch(f, 1, p->precision, eng, p->schedule));
    if (p->chunk <= 0) r.chunk = b->preferred_chunk();
    std::vector<Shard> sh(1);
    sh[0].b = b.get(); sh[0].dev = f->device; sh[0].off = 0; sh[0].R = 1;
    drive(sh, r, ODESAT_MODE_BATCH, false, [&](size_t) { upload_host<TH>(*b, v, xs, xl, true); });
    int64_t solved = -1;
    b->status(&solved, nullptr);
    download_host<TH>(*b, v, xs, xl);
    if (assignment) b->assignment(0, assignment);                                // system.rs:238
    if (steps_taken) *steps_taken = solved >= 0 ? solved + 1 : (r.steps < 0 ? b->step : r.steps);
    if (allsat) *allsat = solved >= 0 ? 1 : 0;
    if (final_dt) { double d = r.dt; if (!r.fixed) b->get_dt(&d); *final_dt = d; }
}

int resolve_gpus(const odesat_params* p) {
    const int want = p->n_gpus > 0 ? p->n_gpus : 1;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess) { cudaGetLastError(); have = 0; }
    if (have <= 0) throw Error(ODESAT_ECUDA, "no CUDA device available (odesat_b200 has no CPU fallback)");
    ODESAT_REQUIRE(want <= have, "params.n_gpus exceeds the number of visible CUDA devices");
    return want;
}

template <typename TH>
void simulate_batch_impl(const odesat_formula* f, int64_t R, TH* v, TH* xs, TH* xl, uint64_t seed,
                         int64_t replica_offset, const odesat_params* p, int mode, int write_back,
                         int64_t* solved_step, uint8_t* verified, int64_t* winner, uint8_t* assignment,
                         int64_t* steps_run) {
    ODESAT_REQUIRE(f != nullptr, "formula is NULL");
    ODESAT_REQUIRE(R >= 0, "negative replica count");
    ODESAT_REQUIRE(mode == ODESAT_MODE_BATCH || mode == ODESAT_MODE_INTER, "unknown mode");
    Resolved r = resolve(f, p);
    if (mode == ODESAT_MODE_BATCH) ODESAT_REQUIRE(r.steps >= 0, "batch needs a step count (main.rs:96-97)");
    int G = resolve_gpus(p);
    if (R == 0) {   // nothing to integrate (the reference's `batch` loop does not run; `inter` would index states[0])
        if (winner) *winner = -1;
        if (steps_run) *steps_run = 0;
        if (assignment) std::memset(assignment, 0, (size_t)f->N);
        return;
    }
    const bool inter_adaptive = mode == ODESAT_MODE_INTER && !r.fixed;   // one dt shared by all replicas: sequential (Q7)
    // adaptive runs: AUTO takes the tile engine only where it has an adaptive kernel (tile_adaptive.cuh); the
    // sequential adaptive `inter` runs on the gather engine
    const int eng = (!r.fixed && p->engine == ODESAT_ENGINE_AUTO) ? (inter_adaptive ? ODESAT_ENGINE_GATHER : ENGINE_AUTO_ADAPTIVE) : p->engine;
    if (inter_adaptive) G = 1;
    G = (int)std::min<int64_t>(G, R);
    int sub = p->sub_batches;
    if (const char* e = std::getenv("ODESAT_SUB_BATCHES")) sub = std::atoi(e);
    // measured on B200 (4096 replicas, N = 10 000, 20 steps per call): 14.9 / 13.4 / 12.9 / 12.6 / 13.3 ms for 1 / 2 / 4 / 8 /
    // 16 sub-batches against 11.2 ms of pure integration → about 512 replicas per sub-batch, at most 8
    if (sub <= 0) sub = (v != nullptr && r.fixed) ? (int)std::max<int64_t>(1, std::min<int64_t>(8, R / G / 512)) : 1;
    if (inter_adaptive) sub = 1;
    sub = (int)std::max<int64_t>(1, std::min<int64_t>(sub, R / G));
    std::vector<Shard> sh = plan_shards(f, R, G, sub);
    cached_shards(f, sh, p->precision, eng, p->schedule);
    for (size_t k = 0; k < sh.size(); ++k) sh[k].b->followed = k + 1 < sh.size() && sh[k + 1].dev == sh[k].dev;
    if (p->chunk <= 0) r.chunk = sh[0].b->preferred_chunk();
    const int64_t N = f->N, M = f->M;
    // main.rs:283-289: whatever the caller does not supply is generated on the device
    auto prepare = [&](size_t k) {
        Shard& s = sh[k];
        if (!(v && xs && xl)) s.b->init(seed, replica_offset + s.off, !v, !xs, !xl, /*finalize=*/!(v || xs || xl));
        if (v || xs || xl)
            upload_host<TH>(*s.b, v ? v + s.off * N : nullptr, xs ? xs + s.off * M : nullptr, xl ? xl + s.off * M : nullptr,
                            /*reset=*/false);
    };
    std::vector<int64_t> solved((size_t)R, -1);
    std::vector<char> enqueued(sh.size(), 0);   // shards whose results are already enqueued behind their final chunk
    DriveResult dr;
    if (inter_adaptive) {
        // adaptive inter: the replicas share ONE dt and step one after the other (system.rs:312-349, quirk Q7)
        DeviceGuard g(sh[0].dev);
        prepare(0);
        sh[0].b->run_inter_adaptive(r.tol, r.zeta, r.steps);
        dr.key = sh[0].b->first_key();
        dr.run = sh[0].b->step;
        if (dr.key != NO_KEY) dr.run = (dr.key >> 32) + 1;
    } else {
        const bool lockstep = mode == ODESAT_MODE_INTER && write_back != 0 && r.fixed;
        // BATCH: a shard's verification is enqueued behind its final chunk, i.e. it overlaps the other shards' integration
        // (INTER masks flags after the winning step and may stop mid-budget: its results are taken after the loop)
        std::function<void(size_t)> early;
        static const bool early_on = [] { const char* e = std::getenv("ODESAT_EARLY_RESULTS"); return !(e && e[0] == '0'); }();
        if (mode == ODESAT_MODE_BATCH && early_on) early = [&](size_t k) { sh[k].b->results_enqueue(); enqueued[k] = 1; };
        dr = drive(sh, r, mode, lockstep, prepare, early);
    }
    std::vector<uint8_t> ver((size_t)R, 0);
    for (size_t k = 0; k < sh.size(); ++k)                                       // cnf.rs:246-264 on every shard, then
        if (!enqueued[k]) { DeviceGuard g(sh[k].dev); sh[k].b->results_enqueue(); }
    for (Shard& s : sh) { DeviceGuard g(s.dev); s.b->results_collect(solved.data() + s.off, ver.data() + s.off); }   // one sync each
    int64_t win = -1, src = 0;
    if (mode == ODESAT_MODE_BATCH) {
        for (int64_t q = 0; q < R; ++q) if (ver[q]) { win = q; break; }          // main.rs:305-307
        src = win >= 0 ? win : R - 1;
    } else {
        if (r.steps == 0) win = 0;                                               // system.rs:274, 353 (Q8)
        else if (dr.key != NO_KEY) win = dr.key & 0xFFFFFFFFll;
        src = win >= 0 ? win : 0;                                                // system.rs:357
        // flags raised after the winning step belong to steps the reference never runs
        if (dr.key != NO_KEY) for (int64_t q = 0; q < R; ++q) if (solved[q] > (dr.key >> 32)) solved[q] = -1;
    }
    if (solved_step) for (int64_t q = 0; q < R; ++q) solved_step[q] = solved[q];
    if (verified) for (int64_t q = 0; q < R; ++q) verified[q] = ver[q];
    if (winner) *winner = win;
    if (assignment) {
        for (Shard& s : sh)
            if (src >= s.off && src < s.off + s.R) { DeviceGuard g(s.dev); s.b->assignment(src - s.off, assignment); }
    }
    if (steps_run) *steps_run = dr.run;
    if (write_back)
        for (Shard& s : sh) {
            DeviceGuard g(s.dev);
            download_host<TH>(*s.b, v ? v + s.off * N : nullptr, xs ? xs + s.off * M : nullptr, xl ? xl + s.off * M : nullptr);
        }
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

void stoch_search_impl(const odesat_formula* f, int64_t R, uint8_t* v, uint64_t* xl, uint64_t seed, int64_t replica_offset,
                       int64_t steps, int chunk, int write_back, int64_t* solved_step, uint8_t* verified, int64_t* winner,
                       uint8_t* assignment, int64_t* steps_run) {
    ODESAT_REQUIRE(f != nullptr, "formula is NULL");
    ODESAT_REQUIRE(R >= 0, "negative replica count");
    require_device();
    if (R == 0) {
        if (winner) *winner = -1;
        if (steps_run) *steps_run = 0;
        if (assignment) std::memset(assignment, 0, (size_t)f->N);
        return;
    }
    DeviceGuard guard(f->device);
    StochBatch b(f, R);
    if (v || xl) b.upload(v, reinterpret_cast<const unsigned long long*>(xl));
    if (chunk <= 0) chunk = 64;
    int64_t key = NO_KEY;
    while (steps < 0 || b.step < steps) {                                        // stoch.rs:94-106
        b.run(seed, replica_offset, steps < 0 ? chunk : std::min<int64_t>(chunk, steps - b.step));
        key = b.first_key();
        if (key != NO_KEY) break;
    }
    std::vector<int64_t> solved((size_t)R, -1);
    b.status(solved.data());
    const int64_t win = key != NO_KEY ? (key & 0xFFFFFFFFll) : -1;
    if (key != NO_KEY) for (int64_t q = 0; q < R; ++q) if (solved[q] > (key >> 32)) solved[q] = -1;
    if (solved_step) for (int64_t q = 0; q < R; ++q) solved_step[q] = solved[q];
    if (verified) b.verify(verified);                                            // cnf.rs:246-264
    if (winner) *winner = win;
    if (assignment) b.assignment(win >= 0 ? win : 0, assignment);                // stoch.rs:109
    if (steps_run) *steps_run = key != NO_KEY ? (key >> 32) + 1 : b.step;
    if (write_back) b.download(v, reinterpret_cast<unsigned long long*>(xl));
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

int odesat_stoch_step(const odesat_formula* f, uint8_t* v, uint64_t* xl, uint64_t seed, int64_t replica, int64_t step_index,
                      int* allsat) {
    return guarded([&] {
        ODESAT_REQUIRE(f && v && xl, "NULL state or formula");
        ODESAT_REQUIRE(step_index >= 0, "negative step index");
        require_device();
        DeviceGuard guard(f->device);
        StochBatch b(f, 1);
        b.upload(v, reinterpret_cast<const unsigned long long*>(xl));
        b.step = step_index;
        b.run(seed, replica, 1);
        int64_t s = -1;
        b.status(&s);
        b.download(v, reinterpret_cast<unsigned long long*>(xl));
        if (allsat) *allsat = s == step_index ? 1 : 0;
    });
}
int odesat_stoch_search(const odesat_formula* f, int64_t R, uint8_t* v, uint64_t* xl, uint64_t seed, int64_t replica_offset,
                        int64_t steps, int32_t chunk, int32_t write_back, int64_t* solved_step, uint8_t* verified,
                        int64_t* winner, uint8_t* assignment, int64_t* steps_run) {
    return guarded([&] {
        stoch_search_impl(f, R, v, xl, seed, replica_offset, steps, chunk, write_back, solved_step, verified, winner, assignment,
                          steps_run);
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
        ODESAT_REQUIRE(f.M >= 1, "tile schedules need at least one clause");   // ragged lengths: loop clauses (tile_ragged.cuh)
        // the same level construction the tile engine uses for a CTA of `threads` threads
        auto lv = schedule == ODESAT_SCHED_EXACT ? build_tile_levels(f, schedule, threads) : build_balanced_levels(f, threads, tile_items_per_level());
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
        ODESAT_REQUIRE(f != nullptr, "formula is NULL");
        DeviceGuard g(f->device);
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
        DeviceGuard g(b->impl->device);
        b->impl->reset();                                    // flags / step / dt of a fresh batch
        b->impl->init(seed, replica_offset, true, true, true);
        b->impl->sync();
    });
}
int odesat_batch_upload(odesat_batch* b, const void* v, const void* xs, const void* xl) {
    return guarded([&] {
        ODESAT_REQUIRE(b && v && xs && xl, "NULL argument");
        DeviceGuard g(b->impl->device);
        b->impl->upload(v, xs, xl, true);
        b->impl->sync();
    });
}
int odesat_batch_download(odesat_batch* b, void* v, void* xs, void* xl) {
    return guarded([&] {
        ODESAT_REQUIRE(b != nullptr, "batch is NULL");
        DeviceGuard g(b->impl->device);
        b->impl->download(v, xs, xl);
    });
}
int odesat_batch_run_fixed(odesat_batch* b, double dt, double zeta, int64_t n, int32_t freeze, float* device_ms) {
    return guarded([&] {
        ODESAT_REQUIRE(b != nullptr, "batch is NULL");
        DeviceGuard g(b->impl->device);
        b->impl->run_fixed(dt, zeta, n, freeze, device_ms);
    });
}
int odesat_batch_run_adaptive(odesat_batch* b, double tolerance, double zeta, int64_t n, float* device_ms) {
    return guarded([&] {
        ODESAT_REQUIRE(b != nullptr, "batch is NULL");
        DeviceGuard g(b->impl->device);
        b->impl->run_adaptive(tolerance, zeta, n, device_ms);
    });
}
int odesat_batch_stream(const odesat_batch* b, void** stream) {
    return guarded([&] {
        ODESAT_REQUIRE(b && stream, "NULL argument");
        *stream = (void*)b->impl->stream;
    });
}
int odesat_batch_run_fixed_async(odesat_batch* b, double dt, double zeta, int64_t n, int32_t freeze, const uint64_t* stop_key) {
    return guarded([&] {
        ODESAT_REQUIRE(b != nullptr, "batch is NULL");
        DeviceGuard g(b->impl->device);
        b->impl->run_fixed_async(dt, zeta, n, freeze, reinterpret_cast<const unsigned long long*>(stop_key));
    });
}
int odesat_batch_post_key(odesat_batch* b, int64_t replica_offset, uint64_t* key_out) {
    return guarded([&] {
        ODESAT_REQUIRE(b && key_out, "NULL argument");
        DeviceGuard g(b->impl->device);
        b->impl->post_key_to(replica_offset, reinterpret_cast<unsigned long long*>(key_out));
    });
}
int odesat_batch_sync(odesat_batch* b) {
    return guarded([&] {
        ODESAT_REQUIRE(b != nullptr, "batch is NULL");
        DeviceGuard g(b->impl->device);
        b->impl->sync();
    });
}
int odesat_batch_status(odesat_batch* b, int64_t* solved_step, int64_t* steps_done) {
    return guarded([&] {
        ODESAT_REQUIRE(b != nullptr, "batch is NULL");
        DeviceGuard g(b->impl->device);
        b->impl->status(solved_step, steps_done);
    });
}
int odesat_batch_first_solved(odesat_batch* b, int64_t* key) {
    return guarded([&] {
        ODESAT_REQUIRE(b && key, "NULL argument");
        DeviceGuard g(b->impl->device);
        *key = b->impl->first_key();
    });
}
int odesat_batch_verify(odesat_batch* b, uint8_t* verified) {
    return guarded([&] {
        ODESAT_REQUIRE(b && verified, "NULL argument");
        DeviceGuard g(b->impl->device);
        b->impl->verify(verified);
    });
}
int odesat_batch_assignment(odesat_batch* b, int64_t replica, uint8_t* assignment) {
    return guarded([&] {
        ODESAT_REQUIRE(b && assignment, "NULL argument");
        DeviceGuard g(b->impl->device);
        b->impl->assignment(replica, assignment);
    });
}
int odesat_batch_dt(odesat_batch* b, double* dt) {
    return guarded([&] {
        ODESAT_REQUIRE(b && dt, "NULL argument");
        DeviceGuard g(b->impl->device);
        b->impl->get_dt(dt);
    });
}

}  // extern "C"
