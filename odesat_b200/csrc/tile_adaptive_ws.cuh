// tile_adaptive_ws.cuh — k_tile_adaptive in the WARP-SPECIALISED, PERSISTENT form of tile_ws.cuh (included by
// tile_engine.cuh).  f32, two replicas per tile, BALANCED schedule with wide levels, fast arithmetic only (the literal
// first step after an import outside the fast domain is taken by k_tile_adaptive<…, STRICT>).
//
// Same algorithm and same arithmetic as k_tile_adaptive (two walks over the schedule per adaptive step: pass A = k1, the
// flag and C_m; pass B = the half / full memories rebuilt from C_m, k2, the new state and the error norm; per-replica dt),
// with the structure of k_tile_ws:
//   * a PRODUCER WARP feeds the ring with bulk copies (cp.async.bulk → full / empty mbarriers): the {xs, xl} cells, the
//     clause words and — for pass B items — the C_m cells that the consumers wrote in pass A (generic-proxy stores, made
//     visible to the bulk reads by the consumers' fence.proxy.async at the start of every item, exactly like the
//     {xs, xl} write-backs);
//   * block barriers only between levels and only among the consumers;
//   * persistent CTAs that take (sub-chunk of steps, tile) work items from the queue, so a shard with few tiles per SM
//     (an adaptive batch sharded over 8 GPUs: 256 tiles on 148 SMs) has no tail wave; a tile's dt travels between
//     sub-chunks through the dt array like v through vt.
// Shared memory: rows[N] (16 B) | ring_m[D][NT] (16 B) | ring_e[D][NT] (8 B) | ring_c[D][NT] (8 B) | items[n_items + 2]
//                | full[D], empty[D] mbarriers | work word | error words
#pragma once

namespace odesat {

template <int NT, int D>
__global__ void __launch_bounds__(NT + 32, 1) k_tile_adaptive_ws(const TileAdaptArgs<float> aa, const TileWork wk) {
    using T = float;
    constexpr int W = 2;
    using Row = float4;
    using Mem = float4;
    using EB = ErrBits<float>;
    using U = unsigned;
    const TileArgs<T>& a = aa.t;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Row* rows = reinterpret_cast<Row*>(smem_raw);
    Mem* ring_m = reinterpret_cast<Mem*>(smem_raw + (size_t)a.N * sizeof(Row));
    uint2* ring_e = reinterpret_cast<uint2*>(ring_m + D * NT);
    uint2* ring_c = ring_e + D * NT;
    uint2* s_items = ring_c + D * NT;
    const int n_items = (a.n_items + D - 1) / D * D;           // walked in whole rings: padded with empty items (≤ 2 extra)
    unsigned long long* full = reinterpret_cast<unsigned long long*>(s_items + a.n_items + 2);
    unsigned long long* empty = full + D;
    volatile int* s_work = reinterpret_cast<volatile int*>(empty + D);
    U* s_err = reinterpret_cast<U*>(const_cast<int*>(s_work) + 2);

    const int s_first = launch_first_step<false>(a);   // block-uniform
    if (s_first >= a.nsteps) return;
    const unsigned tid = threadIdx.x;
    const unsigned lane = tid & 31u;
    const bool producer = tid >= NT;
    const uint2* entries = reinterpret_cast<const uint2*>(a.entry);

    for (int i = tid; i < n_items; i += NT + 32) {
        const uint32_t it = i < a.n_items ? a.items[i] : 0u;
        s_items[i] = make_uint2(it & 0xFFFFFu, ((it >> 20) & 0x7FFu) | (it & TILE_ITEM_LAST));
    }
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < D; ++k) { mbar_init(full + k, 1); mbar_init(empty + k, NT / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_err[0] = EB::NONE;
        s_err[1] = EB::NONE;
    }
    unsigned fpar = 0, epar = (1u << D) - 1u;
    const int total = wk.tiles * wk.nsub;
    const float hi_s = 1.0f - Kc<float>::EPSILON;

    for (;;) {
        __syncthreads();
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
            uint2* vfull = reinterpret_cast<uint2*>(aa.vfull) + tile * a.N;
            Mem* tile_mem = a.mem + tile * a.Mpad;
            uint2* tile_cm = reinterpret_cast<uint2*>(aa.cm) + tile * a.Mpad;
            bool valid[W], frozen[W];
            int32_t solved_at[W];
            T dtw[W];
#pragma unroll
            for (int w = 0; w < W; ++w) {
                valid[w] = tile * W + w < a.R;
                solved_at[w] = valid[w] ? __ldcg(a.solved + tile * W + w) : 0;
                frozen[w] = !valid[w] || solved_at[w] >= 0;
                dtw[w] = valid[w] ? __ldcg(aa.dt + tile * W + w) : 0.01f;
            }
            auto refill = [&](int k, uint2 it, bool with_cm) {   // producer lane 0 only
                const unsigned cnt = it.y & 0x7FFFFFFFu;
                if (cnt == 0) return;
                mbar_wait(empty + k, (epar >> k) & 1u);
                epar ^= 1u << k;
                const unsigned bm = cnt * (unsigned)sizeof(Mem), be = ((cnt + 1u) & ~1u) * 8u;
                mbar_expect_tx(full + k, bm + be + (with_cm ? be : 0u));
                bulk_g2s(ring_m + k * NT, tile_mem + it.x, bm, full + k);
                bulk_g2s(ring_e + k * NT, entries + it.x, be, full + k);
                if (with_cm) bulk_g2s(ring_c + k * NT, tile_cm + it.x, be, full + k);
            };
            if (producer) {
                if (tid == NT) {
#pragma unroll
                    for (int k = 0; k < D; ++k) refill(k, s_items[k], false);
                }
            } else {
                for (int i = tid; i < a.N; i += NT) {
                    const float2 x = __ldcg(reinterpret_cast<const float2*>(vt) + i);
                    rows[i] = make_float4(x.x, x.y, 0.0f, 0.0f);
                }
            }
            __syncthreads();

            for (int s = sa; s < sb; ++s) {
                bool all_frozen = true;
#pragma unroll
                for (int w = 0; w < W; ++w) all_frozen = all_frozen && frozen[w];
                if (all_frozen) break;
                float mx[2] = {0.0f, 0.0f};
                U e_loc[W] = {EB::NONE, EB::NONE};
                const float2 dt2 = make_float2(dtw[0], dtw[1]);
                const float2 h2 = make_float2(0.5f * dtw[0], 0.5f * dtw[1]);                 // :128
                for (int pass = 0; pass < 2; ++pass) {
                    if (producer) {
                        if (tid == NT) {
                            for (int base = 0; base < n_items; base += D) {
#pragma unroll
                                for (int k = 0; k < D; ++k) {
                                    int nx = base + k + D;       // stage k: item base + k → item base + k + D, wrapping into the other pass
                                    bool nb = pass != 0;
                                    if (nx >= n_items) { nx -= n_items; nb = !nb; }
                                    refill(k, s_items[nx], nb);
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
                                // generic-proxy stores of earlier items ({xs, xl} write-backs, C_m) → bulk reads of the same slots
                                asm volatile("fence.proxy.async.global;" ::: "memory");
                                mbar_wait(full + k, (fpar >> k) & 1u);
                                fpar ^= 1u << k;
                                if (tid < cnt) {
                                    const Mem mm = ring_m[k * NT + tid];
                                    const uint2 e = ring_e[k * NT + tid];
                                    Row* const r0 = reinterpret_cast<Row*>(smem_raw + (e.x & 0x3FFF0u));
                                    Row* const r1 = reinterpret_cast<Row*>(smem_raw + ((e.x >> 14) & 0x3FFF0u));
                                    Row* const r2 = reinterpret_cast<Row*>(smem_raw + (e.y & 0x3FFF0u));
                                    const float qf[3] = {(e.y >> 24) & 1u ? -1.0f : 1.0f, (e.y >> 25) & 1u ? -1.0f : 1.0f, (e.y >> 26) & 1u ? -1.0f : 1.0f};
                                    const float4 a0 = *r0, a1 = *r1, a2 = *r2;
                                    const float2 v2[3] = {make_float2(a0.x, a0.y), make_float2(a1.x, a1.y), make_float2(a2.x, a2.y)};
                                    float2 d2[3] = {make_float2(a0.z, a0.w), make_float2(a1.z, a1.w), make_float2(a2.z, a2.w)};
                                    const float2 xs2 = make_float2(mm.x, mm.y), xl2 = make_float2(mm.z, mm.w);
                                    if (pass == 0) {
                                        const float2 mn = clause_rhs_f32x2(v2, d2, qf, xs2, xl2);
                                        mx[0] = rmax(mx[0], mn.x);              // :88 as a running maximum of the clause minima
                                        mx[1] = rmax(mx[1], mn.y);
                                        const float2 cm = mul2(bc2(0.5f), mn);                                            // :60
                                        tile_cm[it.x + tid] = make_uint2(__float_as_uint(cm.x), __float_as_uint(cm.y));
                                    } else {
                                        const uint2 cu = ring_c[k * NT + tid];
                                        const float2 cm1 = make_float2(__uint_as_float(cu.x), __uint_as_float(cu.y));
                                        float2 dxs1, dxl1, dxs2, dxl2;
                                        mem_derivs_f32x2(xs2, cm1, dxs1, dxl1);                                           // :84-85
                                        const float2 xs_f = euler_clamp_f32x2(xs2, dxs1, dt2, Kc<float>::EPSILON, hi_s);  // :125
                                        const float2 xl_f = euler_clamp_f32x2(xl2, dxl1, dt2, 1.0f, a.xl_max);
                                        const float2 xs_h = euler_clamp_f32x2(xs2, dxs1, h2, Kc<float>::EPSILON, hi_s);   // :128
                                        const float2 xl_h = euler_clamp_f32x2(xl2, dxl1, h2, 1.0f, a.xl_max);
                                        const float2 mn = clause_rhs_f32x2(v2, d2, qf, xs_h, xl_h);                       // :129
                                        mem_derivs_f32x2(xs_h, mul2(bc2(0.5f), mn), dxs2, dxl2);
                                        const float2 xs_n = euler_clamp_f32x2(xs_h, dxs2, h2, Kc<float>::EPSILON, hi_s);  // :130
                                        const float2 xl_n = euler_clamp_f32x2(xl_h, dxl2, h2, 1.0f, a.xl_max);
                                        const float es[2] = {fabsf(__fsub_rn(xs_f.x, xs_n.x)), fabsf(__fsub_rn(xs_f.y, xs_n.y))};   // :104-107
                                        const float el[2] = {fabsf(__fsub_rn(xl_f.x, xl_n.x)), fabsf(__fsub_rn(xl_f.y, xl_n.y))};
                                        float xs[2] = {mm.x, mm.y}, xl[2] = {mm.z, mm.w};
                                        const float ns[2] = {xs_n.x, xs_n.y}, nl[2] = {xl_n.x, xl_n.y};
#pragma unroll
                                        for (int w = 0; w < W; ++w) {
                                            if (!frozen[w]) {
                                                if (es[w] == es[w]) { const U b = EB::enc(es[w]); e_loc[w] = b > e_loc[w] ? b : e_loc[w]; }
                                                if (el[w] == el[w]) { const U b = EB::enc(el[w]); e_loc[w] = b > e_loc[w] ? b : e_loc[w]; }
                                                xs[w] = ns[w];
                                                xl[w] = nl[w];
                                            }
                                        }
                                        __stcg(tile_mem + it.x + tid, make_float4(xs[0], xs[1], xl[0], xl[1]));
                                    }
                                    reinterpret_cast<float2*>(r0)[1] = d2[0];
                                    reinterpret_cast<float2*>(r1)[1] = d2[1];
                                    reinterpret_cast<float2*>(r2)[1] = d2[2];
                                }
                                __syncwarp();
                                if (lane == 0) mbar_arrive(empty + k);
                                if ((int)it.y < 0) named_bar_sync(1, NT);     // last item of a level
                            }
                        }
                    }
                    if (pass == 0) {
                        // -------------------- flag (:120-122) + variable pass A ---------------------------
                        const bool u0 = !producer && !(mx[0] < 0.5f), u1 = !producer && !(mx[1] < 0.5f);
                        const unsigned any_unsat = (__syncthreads_or((int)u0) ? 1u : 0u) | (__syncthreads_or((int)u1) ? 2u : 0u);
#pragma unroll
                        for (int w = 0; w < W; ++w) {
                            if (!frozen[w] && !((any_unsat >> w) & 1u)) {   // all satisfied: state untouched
                                solved_at[w] = a.step0 + s;
                                if (tid == 0) __stcg(a.solved + tile * W + w, solved_at[w]);
                                frozen[w] = true;
                            }
                        }
                        if (!producer) {
                            for (int i = tid; i < a.N; i += NT) {
                                const float4 r = rows[i];
                                float v[2] = {r.x, r.y}, vf[2];
                                const float dv[2] = {r.z, r.w};
#pragma unroll
                                for (int w = 0; w < W; ++w) {
                                    vf[w] = frozen[w] ? v[w] : euler_clamp(v[w], dv[w], dtw[w], -1.0f, 1.0f);        // :125
                                    v[w] = frozen[w] ? v[w] : euler_clamp(v[w], dv[w], 0.5f * dtw[w], -1.0f, 1.0f);  // :128
                                }
                                __stcg(vfull + i, make_uint2(__float_as_uint(vf[0]), __float_as_uint(vf[1])));
                                rows[i] = make_float4(v[0], v[1], 0.0f, 0.0f);
                            }
                        }
                        __syncthreads();
                    } else {
                        // -------------------- variable pass B (:130) + error norm + dt (:132-135) ---------
                        if (!producer) {
                            for (int i0 = tid; i0 < a.N; i0 += 4 * NT) {
                                uint2 vfu[4];
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const int i = i0 + u * NT;
                                    vfu[u] = i < a.N ? __ldcg(vfull + i) : make_uint2(0u, 0u);
                                }
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const int i = i0 + u * NT;
                                    if (i < a.N) {
                                        const float4 r = rows[i];
                                        float v[2] = {r.x, r.y};
                                        const float dv[2] = {r.z, r.w}, vf[2] = {__uint_as_float(vfu[u].x), __uint_as_float(vfu[u].y)};
#pragma unroll
                                        for (int w = 0; w < W; ++w) {
                                            if (!frozen[w]) {
                                                v[w] = euler_clamp(v[w], dv[w], 0.5f * dtw[w], -1.0f, 1.0f);
                                                const float e = fabsf(vf[w] - v[w]);                                 // :102-103
                                                if (e == e) { const U b = EB::enc(e); e_loc[w] = b > e_loc[w] ? b : e_loc[w]; }
                                            }
                                        }
                                        rows[i] = make_float4(v[0], v[1], 0.0f, 0.0f);
                                    }
                                }
                            }
#pragma unroll
                            for (int w = 0; w < W; ++w) {
                                const U m = __reduce_max_sync(0xFFFFFFFFu, e_loc[w]);
                                if (lane == 0 && m != EB::NONE) atomicMax(&s_err[w], m);
                            }
                        }
                        __syncthreads();
#pragma unroll
                        for (int w = 0; w < W; ++w) {
                            if (!frozen[w]) {
                                const float e = EB::dec(s_err[w]);
                                dtw[w] = rmax(rmin(dtw[w] * sqrtf(aa.tol / e), 1e3f), 0.0078125f);                   // :133-135
                            }
                        }
                        __syncthreads();
                        if (tid < W) s_err[tid] = EB::NONE;
                    }
                }
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
                for (int i = tid; i < a.N; i += NT) {
                    const float4 r = rows[i];
                    __stcg(reinterpret_cast<float2*>(vt) + i, make_float2(r.x, r.y));
                }
                if (tid == 0) {
#pragma unroll
                    for (int w = 0; w < W; ++w) if (valid[w]) __stcg(aa.dt + tile * W + w, dtw[w]);
                }
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
