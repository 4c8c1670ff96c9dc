// stoch.cuh — the reference's stochastic local search (src/stoch.rs:26-110) for a batch of independent replicas.
//
// A different algorithm from the ODE integrator (integer clause weights, random flips), but the same machinery:
// clause CSR + variable→clause transpose, replica-fastest layouts, one clause phase and one variable phase per step,
// per-replica flags with freezing, chunked early-exit polling.  SURVEY.md §8f row 4.
//
//   step (stoch.rs:26-78), per replica:
//     clause m:   sat = any literal true;  xl_m ← sat ? max(xl_m ⊖ 1, 1) : xl_m ⊕ 20   (⊖ / ⊕ saturating, u64)
//     variable i: total = Σ xl_m over the occurrences of i,  unsat = Σ over those in unsatisfied clauses (UPDATED xl);
//                 r uniform in 1..=total;  flip v_i iff r <= unsat
//     returns "all clauses satisfied" (of the state BEFORE the flips; then unsat = 0 everywhere and nothing flips)
//
// Layouts: v as BIT PLANES  vbits[N][Rw] (bit b of word w = replica 32·w + b): the satisfaction of a clause for 32
// replicas is three word loads and a few logic ops, identical for the whole warp; xl[M][Rp] u64 and the clause
// satisfaction words sat[M][Rw], replica fastest.  The reference draws from an OS-seeded ThreadRng (unreproducible
// by construction); here r = 1 + mulhi64(SplitMix64(seed, replica, step, variable), total) — the same function in
// the oracle, so the GPU trajectories are bit-identical to the oracle's.
#pragma once
#include <algorithm>
#include <limits>
#include <vector>

#include "common.cuh"
#include "formula.hpp"

namespace odesat {

constexpr unsigned long long STOCH_ALPHA = 20ull;   // stoch.rs:18

__host__ __device__ inline uint64_t stoch_bits(uint64_t seed, uint64_t replica, uint64_t step, uint64_t var) {
    return sm64(v0_key(seed, replica) ^ sm64(step + 0x632BE59BD9B4E019ull) ^ (var * 0xD1342543DE82EF95ull));
}

struct StochArgs {
    FormulaDev f;
    int64_t R = 0, Rw = 0;             // replicas, 32-replica words
    uint32_t* vbits = nullptr;         // [N][Rw]
    unsigned long long* xl = nullptr;  // [M][Rw·32]
    uint32_t* sat = nullptr;           // [M][Rw]
    int32_t* solved = nullptr;         // [Rw·32] first flagged step, -1 = none
    uint32_t* unsat_any = nullptr;     // [Rw·32]
    uint64_t seed = 0;
    int64_t replica_offset = 0;
    int32_t step = 0;
};

// one warp = one (clause, replica word): block (32, 8), grid (ceil(M / 8), Rw)
__global__ void __launch_bounds__(256) k_stoch_clause(const StochArgs a) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    const int64_t w = blockIdx.y;
    if (m >= a.f.M) return;
    const unsigned lane = threadIdx.x;
    const int64_t r = w * 32 + lane;
    const bool active = r < a.R && a.solved[r] < 0;
    uint32_t sat_word = 0u;
    for (int j = a.f.coff[m]; j < a.f.coff[m + 1]; ++j) {          // stoch.rs:20-25 evaluate_clause
        const int lit = a.f.lits[j];
        const int var = (lit < 0 ? -lit : lit) - 1;
        sat_word |= a.vbits[(int64_t)var * a.Rw + w] ^ (lit < 0 ? 0xFFFFFFFFu : 0u);
    }
    const bool sat = (sat_word >> lane) & 1u;
    if (active) {
        const int64_t at = m * (a.Rw * 32) + r;
        unsigned long long x = a.xl[at];
        if (sat) { x = x > 0 ? x - 1 : 0; x = x < 1 ? 1 : x; }                       // :48 saturating_sub(1).max(1)
        else { x = x > ~0ull - STOCH_ALPHA ? ~0ull : x + STOCH_ALPHA; }              // :50 saturating_add(ALPHA)
        a.xl[at] = x;
        if (!sat && a.unsat_any[r] == 0u) a.unsat_any[r] = 1u;                       // :61-63
    }
    if (lane == 0) a.sat[m * a.Rw + w] = sat_word;
}

// one warp = one (variable, replica word): block (32, 8), grid (ceil(N / 8), Rw).  Row 0 commits the flags.
__global__ void __launch_bounds__(256) k_stoch_var(const StochArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    const int64_t w = blockIdx.y;
    if (i >= a.f.N) return;
    const unsigned lane = threadIdx.x;
    const int64_t r = w * 32 + lane;
    const bool active = r < a.R && a.solved[r] < 0;
    unsigned long long total = 0, unsat = 0;
    for (int e = a.f.voff[i]; e < a.f.voff[i + 1]; ++e) {                            // :54-59 slab sums, in clause order
        const int m = a.f.occ_clause[e];
        const uint32_t sw = a.sat[(int64_t)m * a.Rw + w];
        if (active) {
            const unsigned long long x = a.xl[(int64_t)m * (a.Rw * 32) + r];
            total += x;
            if (!((sw >> lane) & 1u)) unsat += x;
        }
    }
    bool flip = false;
    if (active && total > 0) {                                                       // :68-73 gen_range(1..=total) <= unsat
        const unsigned long long rnd = 1ull + __umul64hi(stoch_bits(a.seed, (uint64_t)(a.replica_offset + r), (uint64_t)a.step, (uint64_t)i), total);
        flip = rnd <= unsat;
    }
    const uint32_t flips = __ballot_sync(0xFFFFFFFFu, flip);
    if (lane == 0 && flips) a.vbits[i * a.Rw + w] ^= flips;
    if (i == 0 && r < a.R) {
        if (a.solved[r] < 0 && a.unsat_any[r] == 0u) a.solved[r] = a.step;           // :77: the step returned true
        a.unsat_any[r] = 0u;
    }
}

// host v[R][N] bytes ↔ bit planes
__global__ void k_stoch_pack(const uint8_t* __restrict__ v, int64_t R, int64_t N, int64_t Rw, uint32_t* __restrict__ vbits) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    const int64_t w = blockIdx.y;
    if (i >= N) return;
    const int64_t r = w * 32 + threadIdx.x;
    const bool bit = r < R && v[r * N + i] != 0;
    const uint32_t word = __ballot_sync(0xFFFFFFFFu, bit);
    if (threadIdx.x == 0) vbits[i * Rw + w] = word;
}
__global__ void k_stoch_unpack(uint8_t* __restrict__ v, int64_t R, int64_t N, int64_t Rw, const uint32_t* __restrict__ vbits) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    const int64_t w = blockIdx.y;
    if (i >= N) return;
    const int64_t r = w * 32 + threadIdx.x;
    if (r < R) v[r * N + i] = (vbits[i * Rw + w] >> threadIdx.x) & 1u;
}
// host xl[R][M] ↔ device xl[M][Rp]
__global__ void k_stoch_xl_in(const unsigned long long* __restrict__ src, int64_t R, int64_t M, int64_t Rp, unsigned long long* __restrict__ dst) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    const int64_t r = (int64_t)blockIdx.y * 32 + threadIdx.x;
    if (m < M && r < Rp) dst[m * Rp + r] = r < R ? src[r * M + m] : 1ull;
}
__global__ void k_stoch_xl_out(unsigned long long* __restrict__ dst, int64_t R, int64_t M, int64_t Rp, const unsigned long long* __restrict__ src) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    const int64_t r = (int64_t)blockIdx.y * 32 + threadIdx.x;
    if (m < M && r < R) dst[r * M + m] = src[m * Rp + r];
}
__global__ void k_stoch_fill(unsigned long long* p, int64_t n, unsigned long long x) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = x;
}
// cnf.rs:246-264 on the bit planes: bad[r] = some clause falsified by replica r's current v
__global__ void __launch_bounds__(256) k_stoch_verify(const FormulaDev f, const uint32_t* __restrict__ vbits, int64_t R, int64_t Rw,
                                                      uint32_t* __restrict__ bad_words) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t w = blockIdx.y;
    if (m >= f.M) return;
    uint32_t sat_word = 0u;
    for (int j = f.coff[m]; j < f.coff[m + 1]; ++j) {
        const int lit = f.lits[j];
        const int var = (lit < 0 ? -lit : lit) - 1;
        sat_word |= vbits[(int64_t)var * Rw + w] ^ (lit < 0 ? 0xFFFFFFFFu : 0u);
    }
    if (~sat_word) atomicOr(bad_words + w, ~sat_word);
}
__global__ void k_stoch_assignment(const uint32_t* __restrict__ vbits, int64_t N, int64_t Rw, int64_t rep, uint8_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) out[i] = (vbits[i * Rw + rep / 32] >> (rep % 32)) & 1u;
}

// Device-resident batch of R searches.
struct StochBatch {
    const odesat_formula* f;
    int64_t R, Rw, Rp;
    int64_t step = 0, launches = 0;
    cudaStream_t stream = nullptr;
    DevBuf<uint32_t> vbits, sat, unsat_any, bad;
    DevBuf<unsigned long long> xl, stage_xl, key;
    DevBuf<uint8_t> stage_v;
    DevBuf<int32_t> solved;

    StochBatch(const odesat_formula* f_, int64_t R_) : f(f_), R(R_) {
        Rw = std::max<int64_t>((R + 31) / 32, 1);
        Rp = Rw * 32;
        ODESAT_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        vbits.alloc((size_t)(std::max<int64_t>(f->N, 1) * Rw));
        sat.alloc((size_t)(std::max<int64_t>(f->M, 1) * Rw));
        xl.alloc((size_t)(std::max<int64_t>(f->M, 1) * Rp));
        unsat_any.alloc((size_t)Rp);
        solved.alloc((size_t)Rp);
        bad.alloc((size_t)Rw);
        key.alloc(2);
        reset();
    }
    ~StochBatch() { if (stream) cudaStreamDestroy(stream); }
    StochBatch(const StochBatch&) = delete;
    StochBatch& operator=(const StochBatch&) = delete;

    // stoch.rs:84-87: v = false, xl = 1
    void reset() {
        step = 0;
        ODESAT_CUDA(cudaMemsetAsync(vbits.p, 0, vbits.bytes(), stream));
        ODESAT_CUDA(cudaMemsetAsync(unsat_any.p, 0, unsat_any.bytes(), stream));
        ODESAT_CUDA(cudaMemsetAsync(solved.p, 0xFF, solved.bytes(), stream));
        k_stoch_fill<<<(unsigned)((xl.n + 255) / 256), 256, 0, stream>>>(xl.p, (int64_t)xl.n, 1ull);
        ++launches;
    }
    dim3 rows_grid(int64_t rows) const { return dim3((unsigned)std::max<int64_t>((rows + 7) / 8, 1), (unsigned)Rw, 1); }

    void upload(const uint8_t* v, const unsigned long long* x) {
        if (v && f->N > 0 && R > 0) {
            stage_v.alloc((size_t)(R * f->N));
            ODESAT_CUDA(cudaMemcpyAsync(stage_v.p, v, (size_t)(R * f->N), cudaMemcpyHostToDevice, stream));
            k_stoch_pack<<<rows_grid(f->N), dim3(32, 8), 0, stream>>>(stage_v.p, R, f->N, Rw, vbits.p);
            ++launches;
        }
        if (x && f->M > 0 && R > 0) {
            stage_xl.alloc((size_t)(R * f->M));
            ODESAT_CUDA(cudaMemcpyAsync(stage_xl.p, x, (size_t)(R * f->M) * 8, cudaMemcpyHostToDevice, stream));
            k_stoch_xl_in<<<rows_grid(f->M), dim3(32, 8), 0, stream>>>(stage_xl.p, R, f->M, Rp, xl.p);
            ++launches;
        }
        ODESAT_CUDA(cudaStreamSynchronize(stream));
        ODESAT_CUDA(cudaGetLastError());
    }
    void download(uint8_t* v, unsigned long long* x) {
        if (v && f->N > 0 && R > 0) {
            stage_v.alloc((size_t)(R * f->N));
            k_stoch_unpack<<<rows_grid(f->N), dim3(32, 8), 0, stream>>>(stage_v.p, R, f->N, Rw, vbits.p);
            ++launches;
            ODESAT_CUDA(cudaMemcpyAsync(v, stage_v.p, (size_t)(R * f->N), cudaMemcpyDeviceToHost, stream));
        }
        if (x && f->M > 0 && R > 0) {
            stage_xl.alloc((size_t)(R * f->M));
            k_stoch_xl_out<<<rows_grid(f->M), dim3(32, 8), 0, stream>>>(stage_xl.p, R, f->M, Rp, xl.p);
            ++launches;
            ODESAT_CUDA(cudaMemcpyAsync(x, stage_xl.p, (size_t)(R * f->M) * 8, cudaMemcpyDeviceToHost, stream));
        }
        ODESAT_CUDA(cudaStreamSynchronize(stream));
        ODESAT_CUDA(cudaGetLastError());
    }

    // n steps of stoch.rs:26-78 on every replica that has not flagged yet (search's `break`, stoch.rs:96-98)
    void run(uint64_t seed, int64_t replica_offset, int64_t n) {
        ODESAT_REQUIRE(step + n < (int64_t(1) << 31) - 4, "step counter overflow");
        if (R == 0) { step += n; return; }
        for (int64_t k = 0; k < n; ++k) {
            StochArgs a;
            a.f = f->dev; a.R = R; a.Rw = Rw;
            a.vbits = vbits.p; a.xl = xl.p; a.sat = sat.p; a.solved = solved.p; a.unsat_any = unsat_any.p;
            a.seed = seed; a.replica_offset = replica_offset; a.step = (int32_t)(step + k);
            if (f->M > 0) { k_stoch_clause<<<rows_grid(f->M), dim3(32, 8), 0, stream>>>(a); ++launches; }
            k_stoch_var<<<rows_grid(std::max<int64_t>(f->N, 1)), dim3(32, 8), 0, stream>>>(a);
            ++launches;
        }
        step += n;
        ODESAT_CUDA(cudaGetLastError());
    }
    // (first flagged step << 32 | replica), INT64_MAX when none — same key as the integrator's early exit
    int64_t first_key() {
        k_first_key<<<1, 1024, 0, stream>>>(solved.p, R, 0, key.p);
        ++launches;
        unsigned long long h[2] = {0, 0};
        ODESAT_CUDA(cudaMemcpyAsync(h, key.p, 16, cudaMemcpyDeviceToHost, stream));
        ODESAT_CUDA(cudaStreamSynchronize(stream));
        return (int64_t)h[0];
    }
    void status(int64_t* out) {
        std::vector<int32_t> h((size_t)std::max<int64_t>(R, 1));
        ODESAT_CUDA(cudaMemcpyAsync(h.data(), solved.p, (size_t)R * 4, cudaMemcpyDeviceToHost, stream));
        ODESAT_CUDA(cudaStreamSynchronize(stream));
        for (int64_t r = 0; r < R; ++r) out[r] = h[r];
    }
    void verify(uint8_t* out) {
        ODESAT_CUDA(cudaMemsetAsync(bad.p, 0, bad.bytes(), stream));
        if (f->M > 0) {
            k_stoch_verify<<<dim3((unsigned)((f->M + 255) / 256), (unsigned)Rw), 256, 0, stream>>>(f->dev, vbits.p, R, Rw, bad.p);
            ++launches;
        }
        std::vector<uint32_t> h((size_t)Rw);
        ODESAT_CUDA(cudaMemcpyAsync(h.data(), bad.p, (size_t)Rw * 4, cudaMemcpyDeviceToHost, stream));
        ODESAT_CUDA(cudaStreamSynchronize(stream));
        for (int64_t r = 0; r < R; ++r) out[r] = ((h[r / 32] >> (r % 32)) & 1u) ? 0 : 1;
    }
    void assignment(int64_t r, uint8_t* out) {
        ODESAT_REQUIRE(r >= 0 && r < R, "replica index out of range");
        if (f->N == 0) return;
        stage_v.alloc((size_t)f->N);
        k_stoch_assignment<<<(unsigned)((f->N + 255) / 256), 256, 0, stream>>>(vbits.p, f->N, Rw, r, stage_v.p);
        ++launches;
        ODESAT_CUDA(cudaMemcpyAsync(out, stage_v.p, (size_t)f->N, cudaMemcpyDeviceToHost, stream));
        ODESAT_CUDA(cudaStreamSynchronize(stream));
    }
};

}  // namespace odesat
