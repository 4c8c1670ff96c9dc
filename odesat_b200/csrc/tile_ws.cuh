// tile_ws.cuh — the WARP-SPECIALISED, PERSISTENT form of the tile kernel (included by tile_engine.cuh).
//
// Same algorithm, same arithmetic and same schedule walk as k_tile_fixed / k_tile_fixed_tma (one thread = one
// clause slot of a replica tile whose {v, dv} rows live in shared memory; system.rs:141-154 fused), with three
// structural changes, each aimed at what ncu showed for k_tile_fixed_tma (barrier stalls 3.5 cycles per issued
// instruction, all 24 warps in phase between 56 block barriers per step, one elected thread refilling the ring
// after it leaves the barrier):
//
//  1. PRODUCER WARP.  The CTA is NT consumer threads + one extra warp whose lane 0 does nothing but feed the ring:
//     for every item it waits on the stage's EMPTY mbarrier (one arrival per consumer warp, made once the warp has
//     its cell and clause word in registers), arms the stage's FULL mbarrier with the byte count and issues the two
//     bulk copies (cp.async.bulk, SASS UBLKCP).  Consumers wait on FULL.  The ring is thereby decoupled from the
//     level barriers: a stage is recycled as soon as its last reader is done, not at the next block barrier.
//  2. BARRIERS ONLY BETWEEN LEVELS, and only among the consumers (bar.sync 1, NT).  With the ring decoupled an item
//     no longer has to end with a barrier, so the BALANCED schedule can use levels of SEVERAL items (tile_schedule:
//     as many colours as the maximum variable degree — 28 levels of two 768-clause items at the headline size
//     instead of 56 levels of one): half the barriers, and inside a level the warps drift apart, so the row
//     gathers of one warp overlap the arithmetic of another.
//  3. PERSISTENT CTAs WITH A WORK QUEUE.  The grid is one CTA per SM; work items are (sub-chunk of steps, tile) in
//     sub-chunk-major order, taken with an atomic counter.  A tile's sub-chunk c may start once its sub-chunk c − 1
//     is published (release/acquire on done[tile]); items are handed out in order, so the predecessor is always
//     held by a running CTA and the wait cannot deadlock.  With many tiles per SM one sub-chunk is the whole launch
//     (a plain dynamic tile scheduler); with few — 4 096 replicas over 8 GPUs leave 256 tiles for 148 SMs, 1.73
//     waves — the launch is cut into sub-chunks of a few steps so that every SM gets the same number of
//     (tile, step) units: the tail wave disappears (the rows of a tile are written back and re-read once per
//     sub-chunk, 2 × 80 KB against 1.4 MB of {xs, xl} traffic per step).
#pragma once

namespace odesat {

__device__ __forceinline__ void mbar_arrive(void* bar) {
    asm volatile("{\n.reg .b64 t;\nmbarrier.arrive.shared::cta.b64 t, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
// arrive that the compiler may schedule freely but not before `dep0` / `dep1` (the values loaded from the stage) exist:
// no memory clobber, so it does not pin the surrounding loads and stores the way a plain asm volatile("..." ::: "memory") does
__device__ __forceinline__ void mbar_arrive_after(void* bar, unsigned dep0, unsigned dep1) {
    asm volatile("{\n.reg .b64 t;\n.reg .b32 u;\nand.b32 u, %1, %2;\nmbarrier.arrive.shared::cta.b64 t, [%0];\n}" ::"r"(smem_u32(bar)), "r"(dep0), "r"(dep1));
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// Shared memory: rows[N] (16 B) | ring_m[D][NT] (16 B) | ring_e[D][NT] (8 B) | items[n_items + 2] (8 B)
//                | full[D], empty[D] mbarriers | work word
// Measured on B200 (headline size, f32, ms per step): releasing the stage at the END of the item 0.5395; releasing it
// between the arithmetic and the stores (the asm statement pins ptxas' schedule there) 0.5596; clause words loaded
// straight into registers one item ahead (ld.global.nc) instead of travelling through the ring 0.61 — both removed; the
// producer also prefetching the cells of a later item into L2 (cp.async.bulk.prefetch.L2, 1–16 items beyond the ring)
// 0.532–0.533 against 0.526 at 704 threads / ring of four — removed.
template <typename T, int NT, int D>
__global__ void __launch_bounds__(NT + 32, 1) k_tile_ws(const TileArgs<T> a, const TileWork wk) {
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
    unsigned long long* full = reinterpret_cast<unsigned long long*>(s_items + n_items + 2);
    unsigned long long* empty = full + D;
    volatile int* s_work = reinterpret_cast<volatile int*>(empty + D);

    const int s_first = launch_first_step<false>(a);   // block-uniform
    if (s_first >= a.nsteps) return;
    const unsigned tid = threadIdx.x;
    const unsigned lane = tid & 31u;
    const bool producer = tid >= NT;                           // the extra warp; its lane 0 feeds the ring
    const uint2* entries = reinterpret_cast<const uint2*>(a.entry);

    for (int i = tid; i < n_items; i += NT + 32) {
        const uint32_t it = a.items[i];
        s_items[i] = make_uint2(it & 0xFFFFFu, ((it >> 20) & 0x7FFu) | (it & TILE_ITEM_LAST));
    }
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < D; ++k) { mbar_init(full + k, 1); mbar_init(empty + k, NT / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // phase bits: consumers wait on FULL (first phase 0), the producer on EMPTY starting with the phase that
    // precedes the first one, which a fresh mbarrier reports as complete — so the first fill of a stage does not wait
    unsigned fpar = 0, epar = (1u << D) - 1u;
    const int total = wk.tiles * wk.nsub;

    for (;;) {
        __syncthreads();                                       // s_work is free (and, first time, the prologue is visible)
        if (tid == 0) *s_work = atomicAdd(wk.counter, 1);
        __syncthreads();
        const int widx = *s_work;
        if (widx >= total) break;
        const int sub = widx / wk.tiles;
        const int64_t tile = widx - sub * wk.tiles;
        const int sa = s_first + sub * wk.ksub, sb = min(a.nsteps, sa + wk.ksub);
        if (sub > 0 && tid == 0) {
            while (ld_acquire_gpu(wk.done + tile) < sub) { }
        }
        __syncthreads();
        if (sa < sb) {
            T* vt = a.vt + tile * a.N * W;
            Mem* tile_mem = a.mem + tile * a.Mpad;
            bool valid[W], frozen[W];
            int32_t solved_at[W];
#pragma unroll
            for (int w = 0; w < W; ++w) {
                valid[w] = tile * W + w < a.R;
                solved_at[w] = valid[w] ? __ldcg(a.solved + tile * W + w) : 0;
                frozen[w] = !valid[w] || (a.freeze && solved_at[w] >= 0);
            }
            auto refill = [&](int k, uint2 it) {   // producer lane 0 only
                const unsigned cnt = it.y & 0x7FFFFFFFu;
                if (cnt == 0) return;
                mbar_wait(empty + k, (epar >> k) & 1u);          // every consumer warp has read the stage's previous item
                epar ^= 1u << k;
                const unsigned bm = cnt * (unsigned)sizeof(Mem), be = ((cnt + 1u) & ~1u) * 8u;
                mbar_expect_tx(full + k, bm + be);
                bulk_g2s(ring_m + k * NT, tile_mem + it.x, bm, full + k);
                bulk_g2s(ring_e + k * NT, entries + it.x, be, full + k);
            };
            if (producer) {                                    // the ring fills while the consumers fetch the rows
                if (tid == NT) {
#pragma unroll
                    for (int k = 0; k < D; ++k) refill(k, s_items[k]);
                }
            } else if constexpr (W == 2) {
                for (int i = tid; i < a.N; i += NT) {
                    const float2 x = __ldcg(reinterpret_cast<const float2*>(vt) + i);
                    rows[i] = make_float4(x.x, x.y, 0.0f, 0.0f);
                }
            } else {
                for (int i = tid; i < a.N; i += NT) {
                    T v[W], dv[W];
#pragma unroll
                    for (int w = 0; w < W; ++w) { v[w] = __ldcg(vt + (int64_t)i * W + w); dv[w] = T(0); }
                    rows[i] = IO::pack(v, dv);
                }
            }
            __syncthreads();                                   // rows are in place

            for (int s = sa; s < sb; ++s) {
                bool all_frozen = true;
#pragma unroll
                for (int w = 0; w < W; ++w) all_frozen = all_frozen && frozen[w];
                if (all_frozen) break;
                bool unsat[W];
                float mx[2] = {0.0f, 0.0f};
                T dtw[W];
#pragma unroll
                for (int w = 0; w < W; ++w) { unsat[w] = false; dtw[w] = frozen[w] ? T(0) : a.dt; }
                if (producer) {
                    if (tid == NT) {
                        for (int base = 0; base < n_items; base += D) {
#pragma unroll
                            for (int k = 0; k < D; ++k) {
                                int nx = base + k + D;           // stage k: item base + k → item base + k + D (wrapping into the next step)
                                if (nx >= n_items) nx -= n_items;
                                refill(k, s_items[nx]);
                            }
                        }
                    }
                } else {
                    for (int base = 0; base < n_items; base += D) {
#pragma unroll
                        for (int k = 0; k < D; ++k) {
                            const uint2 it = s_items[base + k];
                            const unsigned cnt = it.y & 0x7FFFFFFFu;     // block-uniform
                            if (cnt == 0) continue;
                            // generic-proxy write-backs of earlier items → async-proxy (bulk) reads of the same slots one
                            // step later: cumulative over this thread's earlier stores, issued one item late
                            asm volatile("fence.proxy.async.global;" ::: "memory");
                            mbar_wait(full + k, (fpar >> k) & 1u);
                            fpar ^= 1u << k;
                            const bool mine = tid < cnt;
                            T d[3][W], xs[W], xl[W];
                            Row *r0 = nullptr, *r1 = nullptr, *r2 = nullptr;
                            if (mine) {
                                const Mem mm = ring_m[k * NT + tid];
                                const uint2 e = ring_e[k * NT + tid];
                                // (a warp is converged here: items are full except the last of a level, whose partial warp
                                // releases late)
                                if (wk.early && lane == 0 && cnt == NT) mbar_arrive_after(empty + k, e.y, *reinterpret_cast<const unsigned*>(&mm));
                                r0 = reinterpret_cast<Row*>(smem_raw + (e.x & 0x3FFF0u));
                                r1 = reinterpret_cast<Row*>(smem_raw + ((e.x >> 14) & 0x3FFF0u));
                                r2 = reinterpret_cast<Row*>(smem_raw + (e.y & 0x3FFF0u));
                                const T q[3] = {(e.y >> 24) & 1u ? T(-1) : T(1), (e.y >> 25) & 1u ? T(-1) : T(1), (e.y >> 26) & 1u ? T(-1) : T(1)};
                                T v[3][W];
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
                            }
                            if (mine) {
                                IO::store_dv(r0, d[0]);
                                IO::store_dv(r1, d[1]);
                                IO::store_dv(r2, d[2]);
                                __stcg(tile_mem + it.x + tid, IO::pack_mem(xs, xl));
                            }
                            // the warp's cells and clause words have long been consumed (the stores above depend on them):
                            // release the stage to the producer
                            if (!(wk.early && cnt == NT)) {
                                __syncwarp();
                                if (lane == 0) mbar_arrive(empty + k);
                            }
                            if ((int)it.y < 0) named_bar_sync(1, NT);     // last item of a level: the level's dv stores are complete
                        }
                    }
                }
                if constexpr (W == 2 && sizeof(T) == 4) { unsat[0] = !(mx[0] < 0.5f); unsat[1] = !(mx[1] < 0.5f); }
                unsigned any_unsat = 0;
#pragma unroll
                for (int w = 0; w < W; ++w) any_unsat |= (__syncthreads_or((int)(!producer && unsat[w])) ? 1u : 0u) << w;
                if (!producer) {
                    for (int i = tid; i < a.N; i += NT) {
                        T v[W], dv[W];
                        IO::unpack(rows[i], v, dv);
#pragma unroll
                        for (int w = 0; w < W; ++w) { v[w] = euler_clamp(v[w], dv[w], dtw[w], T(-1), T(1)); dv[w] = T(0); }
                        rows[i] = IO::pack(v, dv);
                    }
                }
#pragma unroll
                for (int w = 0; w < W; ++w) {
                    if (valid[w] && !frozen[w] && !((any_unsat >> w) & 1u)) {
                        if (solved_at[w] < 0) {
                            solved_at[w] = a.step0 + s;
                            if (tid == 0) __stcg(a.solved + tile * W + w, solved_at[w]);
                        }
                        if (a.freeze) frozen[w] = true;
                    }
                }
                __syncthreads();
            }
            // the copies requested for a step that does not run: take them and hand the stages back
            if (!producer) {
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    if ((s_items[k].y & 0x7FFFFFFFu) == 0u) continue;
                    mbar_wait(full + k, (fpar >> k) & 1u);
                    fpar ^= 1u << k;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty + k);
                }
                if constexpr (W == 2) {
                    for (int i = tid; i < a.N; i += NT) {
                        const float4 r = rows[i];
                        __stcg(reinterpret_cast<float2*>(vt) + i, make_float2(r.x, r.y));
                    }
                } else {
                    for (int i = tid; i < a.N; i += NT) {
                        T v[W], dv[W];
                        IO::unpack(rows[i], v, dv);
#pragma unroll
                        for (int w = 0; w < W; ++w) __stcg(vt + (int64_t)i * W + w, v[w]);
                    }
                }
                // a later sub-chunk of this tile (any CTA) reads the {xs, xl} slots with bulk copies
                asm volatile("fence.proxy.async.global;" ::: "memory");
            }
        }
        if (wk.nsub > 1) {
            __threadfence();
            __syncthreads();
            if (tid == 0) st_release_gpu(wk.done + tile, sub + 1);
        }
    }
}

}  // namespace odesat
