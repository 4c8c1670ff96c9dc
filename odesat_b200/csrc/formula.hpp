// formula.hpp — host-side formula object: validation, CSR upload, variable→clause transpose.
//
// Boundary type: the reference's `CNFFormula { clauses, varnum }` (cnf.rs:53-57) arrives
// flattened as (varnum, clause_off[M+1], lits[L]).  The transpose lists, for every variable,
// its occurrences sorted by (clause, literal position) — the order in which the reference's
// sequential clause loop adds into dy.v[i] (system.rs:35-80).
#pragma once
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"
#include "host_util.hpp"

namespace odesat {

struct TileSchedule;   // tile_schedule.hpp
struct TileLevels;

}  // namespace odesat

struct odesat_formula {
    int64_t N = 0, M = 0, L = 0;
    int K = 0;                       // uniform clause length (0 = ragged or M == 0)
    bool distinct_vars = true;       // no variable repeated inside a clause
    int max_degree = 0;              // max occurrences of a variable
    int64_t n_loopy = 0;             // clauses with no literal, more than three, or a repeated variable (tile engine: loop clauses)
    int device = 0;
    std::vector<int64_t> h_off;      // [M+1]
    std::vector<int32_t> h_lits;     // [L]
    std::vector<int32_t> h_voff, h_occ_clause, h_occ_slot;
    std::vector<int8_t> h_xs0;
    odesat::DevBuf<int32_t> d_coff, d_lits, d_voff, d_occ_clause, d_occ_slot;
    odesat::DevBuf<int8_t> d_xs0;
    odesat::FormulaDev dev;
    // tile-engine schedules, built on first use: levels keyed by schedule kind, padded
    // schedules keyed by kind * 4096 + warps per CTA
    mutable std::map<int, std::shared_ptr<odesat::TileLevels>> tile_levels;
    mutable std::map<int, std::shared_ptr<odesat::TileSchedule>> tile_sched;
    // copies of this formula on the other CUDA devices of the process (multi-GPU calls), built on first use
    mutable std::map<int, std::unique_ptr<odesat_formula>> peers;
    // device buffers of the last odesat_simulate* call — one batch per (device, sub-batch) shard — reused by the
    // next call of the same shape (allocating and freeing several GB per call costs more than the upload).
    // Declared after `peers`: the batches are destroyed before the formulas they point to.
    mutable std::vector<std::shared_ptr<void>> batch_cache;
    mutable std::string cache_key;

    // SORTED VIEW (large single instances on the gather engine): the same formula with its clauses STORED in the order of
    // their smallest variable, so that neighbouring clause rows gather neighbouring v rows (one of the three random L2
    // sectors per clause becomes a streamed one) and their first contributions land next to each other.  The
    // variable→clause lists keep the ORIGINAL (clause, position) order — the reference's summation order — and only
    // point into the permuted storage.  cperm[p] = original index of the clause stored at p (xs / xl rows are permuted
    // on upload and download).  Built on first use, on the formula's own device.
    struct SortedView {
        odesat::DevBuf<int32_t> coff, lits, occ_clause, occ_slot, cperm;
        odesat::DevBuf<int8_t> xs0;
        odesat::FormulaDev dev;
    };
    mutable std::unique_ptr<SortedView> sorted;
    const SortedView& sorted_view() const {
        using namespace odesat;
        if (sorted) return *sorted;
        std::unique_ptr<SortedView> v(new SortedView);
        std::vector<int32_t> order(M), key(M), pos(M);
        for (int64_t m = 0; m < M; ++m) {
            int32_t k = INT32_MAX;
            for (int64_t j = h_off[m]; j < h_off[m + 1]; ++j) k = std::min(k, std::abs(h_lits[j]));
            key[m] = k;
            order[m] = (int32_t)m;
        }
        std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return key[a] < key[b]; });
        std::vector<int32_t> coff(M + 1, 0), lits(L), occ_c(L), occ_s(L);
        std::vector<int8_t> xs0(M);
        for (int64_t p = 0; p < M; ++p) {
            const int32_t m = order[p];
            pos[m] = (int32_t)p;
            const int64_t len = h_off[m + 1] - h_off[m];
            for (int64_t j = 0; j < len; ++j) lits[coff[p] + j] = h_lits[h_off[m] + j];
            coff[p + 1] = coff[p] + (int32_t)len;
            xs0[p] = h_xs0[m];
        }
        for (int64_t e = 0; e < L; ++e) {
            const int32_t m = h_occ_clause[e];
            occ_c[e] = pos[m];
            occ_s[e] = coff[pos[m]] + (int32_t)(h_occ_slot[e] - h_off[m]);
        }
        auto up = [](auto& buf, const auto& host) {
            buf.alloc(host.size());
            if (!host.empty())
                ODESAT_CUDA(cudaMemcpy(buf.p, host.data(), host.size() * sizeof(host[0]), cudaMemcpyHostToDevice));
        };
        up(v->coff, coff); up(v->lits, lits); up(v->occ_clause, occ_c); up(v->occ_slot, occ_s); up(v->cperm, order); up(v->xs0, xs0);
        v->dev = dev;
        v->dev.coff = v->coff.p; v->dev.lits = v->lits.p; v->dev.occ_clause = v->occ_clause.p; v->dev.occ_slot = v->occ_slot.p;
        v->dev.xs0 = v->xs0.p;
        sorted = std::move(v);
        return *sorted;
    }

    // this formula on CUDA device `dev_id` (the handle itself on its own device)
    const odesat_formula* on_device(int dev_id) const {
        if (dev_id == device) return this;
        auto it = peers.find(dev_id);
        if (it == peers.end()) {
            std::unique_ptr<odesat_formula> c(new odesat_formula);
            c->build(N, M, h_off.data(), h_lits.data());
            int prev = 0;
            ODESAT_CUDA(cudaGetDevice(&prev));
            ODESAT_CUDA(cudaSetDevice(dev_id));
            try { c->upload(); } catch (...) { cudaSetDevice(prev); throw; }
            ODESAT_CUDA(cudaSetDevice(prev));
            it = peers.emplace(dev_id, std::move(c)).first;
        }
        return it->second.get();
    }

    double default_zeta() const {   // system.rs:164-173
        const double d = double(M) / double(N);
        return d >= 6.0 ? 0.1 : (d >= 4.9 ? 0.01 : 0.001);
    }

    void build(int64_t varnum, int64_t n_clauses, const int64_t* off, const int32_t* lits) {
        using namespace odesat;
        ODESAT_REQUIRE(varnum >= 0 && n_clauses >= 0, "negative varnum / clause count");
        ODESAT_REQUIRE(off != nullptr, "clause_off is NULL");
        ODESAT_REQUIRE(off[0] == 0, "clause_off[0] must be 0");
        for (int64_t m = 0; m < n_clauses; ++m)
            ODESAT_REQUIRE(off[m + 1] >= off[m], "clause_off must be non-decreasing");
        N = varnum;
        M = n_clauses;
        L = off[M];
        ODESAT_REQUIRE(L < (int64_t(1) << 31) - 64 && N < (int64_t(1) << 31) - 64 && M < (int64_t(1) << 31) - 64,
                       "formula too large for 32-bit indices");
        ODESAT_REQUIRE(L == 0 || lits != nullptr, "lits is NULL");
        h_off.assign(off, off + M + 1);
        h_lits.assign(lits, lits + L);
        // validate literals (the reference panics on an out-of-bounds index, system.rs:48)
        std::vector<int32_t> deg(N + 1, 0);
        for (int64_t j = 0; j < L; ++j) {
            const int64_t a = h_lits[j] < 0 ? -(int64_t)h_lits[j] : (int64_t)h_lits[j];
            ODESAT_REQUIRE(a >= 1 && a <= N, "literal index outside 1..varnum");
            deg[a - 1]++;
        }
        K = 0;
        if (M > 0) {
            const int64_t k0 = h_off[1] - h_off[0];
            bool uni = k0 > 0 && k0 <= 8;
            for (int64_t m = 0; m < M && uni; ++m) uni = (h_off[m + 1] - h_off[m]) == k0;
            K = uni ? (int)k0 : 0;
        }
        // transpose by counting sort; scanning literals in (clause, position) order keeps
        // every variable's list sorted the way the reference accumulates.
        h_voff.assign(N + 1, 0);
        max_degree = 0;
        for (int64_t i = 0; i < N; ++i) {
            h_voff[i + 1] = h_voff[i] + deg[i];
            if (deg[i] > max_degree) max_degree = deg[i];
        }
        h_occ_clause.resize(L);
        h_occ_slot.resize(L);
        std::vector<int32_t> cur(h_voff.begin(), h_voff.end() - 1);
        h_xs0.resize(M);
        distinct_vars = true;
        n_loopy = 0;
        for (int64_t m = 0; m < M; ++m) {
            bool anyneg = false;
            const bool was_distinct = distinct_vars;
            distinct_vars = true;
            for (int64_t j = h_off[m]; j < h_off[m + 1]; ++j) {
                const int32_t l = h_lits[j];
                const int32_t var = (l < 0 ? -l : l) - 1;
                anyneg = anyneg || l < 0;
                const int32_t e = cur[var]++;
                if (e > h_voff[var] && h_occ_clause[e - 1] == (int32_t)m) distinct_vars = false;
                h_occ_clause[e] = (int32_t)m;
                h_occ_slot[e] = (int32_t)j;
            }
            h_xs0[m] = anyneg ? 1 : -1;   // system.rs:362-372
            const int64_t len = h_off[m + 1] - h_off[m];
            if (len == 0 || len > 3 || !distinct_vars) ++n_loopy;
            distinct_vars = distinct_vars && was_distinct;
        }
    }

    void upload() {
        using namespace odesat;
        ODESAT_CUDA(cudaGetDevice(&device));
        std::vector<int32_t> coff32(M + 1);
        for (int64_t m = 0; m <= M; ++m) coff32[m] = (int32_t)h_off[m];
        auto up = [](auto& buf, const auto& host) {
            buf.alloc(host.size());
            if (!host.empty())
                ODESAT_CUDA(cudaMemcpy(buf.p, host.data(), host.size() * sizeof(host[0]), cudaMemcpyHostToDevice));
        };
        up(d_coff, coff32);
        up(d_lits, h_lits);
        up(d_voff, h_voff);
        up(d_occ_clause, h_occ_clause);
        up(d_occ_slot, h_occ_slot);
        up(d_xs0, h_xs0);
        dev.N = N; dev.M = M; dev.L = L; dev.K = K;
        dev.coff = d_coff.p; dev.lits = d_lits.p; dev.voff = d_voff.p;
        dev.occ_clause = d_occ_clause.p; dev.occ_slot = d_occ_slot.p; dev.xs0 = d_xs0.p;
    }
};
