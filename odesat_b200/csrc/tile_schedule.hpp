// tile_schedule.hpp — host-side "schedule compiler" of the TILE engine.
//
// The tile kernel keeps {v, dv} of a replica tile in shared memory and lets one thread per
// clause add its contributions into dv with a plain read-modify-write.  That is race-free iff
// the clauses processed between two block barriers touch pairwise distinct variables, so the
// formula is compiled once, on the host, into LEVELS (colour classes of the clause-conflict
// graph):
//   EXACT    level(m) = 1 + max over m's variables of the level of the previous clause that
//            contains the variable.  Every variable then meets its clauses in ascending
//            clause index as the levels advance — the order in which the reference's
//            sequential loop executes `dy.v[i] += …` (system.rs:35-80) — so the running sum
//            is bit-identical to the reference's.
//   BALANCED greedy least-loaded colouring into equal-size classes (fewer barriers, no
//            ordering guarantee; results agree to rounding and are run-to-run deterministic).
// Inside a level the clause order is free; it is chosen so that the 8 lanes of a quarter-warp
// (one 128-byte shared-memory wavefront of 16-byte rows) hit 8 distinct bank groups wherever
// possible, and literal positions inside a clause are permuted for the same purpose (the
// per-clause arithmetic is symmetric in the literals when the variables are distinct).
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <queue>
#include <vector>

#include "formula.hpp"

namespace odesat {

// An ITEM is up to NT consecutive slots of one level, processed by the CTA in lock-step (thread
// t takes slot base + t when t < nvalid).  Packed as base (20 bits) | nvalid (11 bits) << 20 |
// last-item-of-level << 31.
constexpr uint32_t TILE_ITEM_LAST = 1u << 31;
inline uint32_t pack_item(uint32_t base, uint32_t nvalid, bool last) {
    return base | (nvalid << 20) | (last ? TILE_ITEM_LAST : 0u);
}

struct TileSchedule {
    int kind = 0;
    int nt = 0;                        // threads per CTA the items were cut for
    int64_t M = 0, Mpad = 0;           // real clauses, slots (levels padded to 8 slots = 128 B)
    int nlev = 0;
    int n_items = 0;
    std::vector<uint32_t> items;       // [n_items] packed (base, nvalid, last) — valid when nt <= 1024 and Mpad < 2^20
    std::vector<uint2> items2;         // [n_items] {base, nvalid | last << 31} — any width (cluster tiles)
    std::vector<int32_t> perm;         // [Mpad] slot → clause index, −1 = padding
    std::vector<uint64_t> entry;       // [Mpad] packed clause: 3×16-bit row + sign bits + valid; or a LOOP clause (below)
    std::vector<uint32_t> aux;         // literal words of the loop clauses: row byte offset | negated << 31, 4-word aligned per clause
    int64_t n_group = 0;               // group clauses: 4..32 literals, distinct variables, one lane per literal
    int64_t n_loop = 0;                // loop clauses: no literal, more than three, or a repeated variable (tile_ragged.cuh)
    double conflict_wavefronts = 0;    // avg shared-memory wavefronts per quarter-warp access (1 = ideal)
    DevBuf<int32_t> d_perm;
    DevBuf<uint32_t> d_items;
    DevBuf<uint2> d_items2;
    DevBuf<uint64_t> d_entry;
    DevBuf<uint32_t> d_aux;
};

// A clause with no literal, more than three, or a repeated variable (ragged formulas — `solve -r` output,
// cnf.rs:397-416; SURVEY quirk Q9's empty clause) is a LOOP clause: its slot's entry holds {offset into TileSchedule::aux, length | TILE_ENTRY_LOOP} and the
// kernel walks its literals (tile_ragged.cuh).  Bit 27 of the high word is clear in every 3-literal entry.
constexpr uint32_t TILE_ENTRY_LOOP = 1u << 27;
// Clauses of one or two literals (distinct variables) keep the packed word: the literal positions that do not exist are
// flagged (TILE_ENTRY_NO2: no third literal, TILE_ENTRY_NO1: no second one either), their row offset is 0 and the kernel
// gives them the value +inf, which is what min / second-min start from (system.rs:46-47).
constexpr uint32_t TILE_ENTRY_NO2 = 1u << 28, TILE_ENTRY_NO1 = 1u << 29;
// GROUP clauses: four to 32 literals with distinct variables are evaluated by as many LANES of one warp, one literal each
// (tile_ragged.cuh): the clause occupies P = 4, 8, 16 or 32 consecutive slots aligned to P inside the warp; slot j holds
// literal j — low word: row byte offset | negated << 31; high word: TILE_ENTRY_GROUP | log2(P) << 8 | j, and
// TILE_ENTRY_VOID for the slots beyond the clause's length.  Only slot 0 (the leader) owns a {xs, xl} cell
// (perm = clause index); the others are padding cells (perm = -1).
constexpr uint32_t TILE_ENTRY_GROUP = 1u << 30, TILE_ENTRY_VOID = 1u << 25;
inline uint64_t pack_entry_group(uint32_t var, bool neg, uint32_t log2p, uint32_t pos, bool live) {
    const uint32_t lo = live ? ((var << 4) | (neg ? 0x80000000u : 0u)) : 0u;
    const uint32_t hi = TILE_ENTRY_GROUP | (log2p << 8) | pos | (live ? 0u : TILE_ENTRY_VOID);
    return (uint64_t)lo | ((uint64_t)hi << 32);
}
// TILE_ENTRY_LOOP8: at most eight literals, all variables distinct — the kernel then takes all rows into registers at once
constexpr uint32_t TILE_ENTRY_LOOP8 = 1u << 26;
inline uint64_t pack_entry_loop(uint32_t aux_off, uint32_t len, bool distinct) {
    return (uint64_t)aux_off | ((uint64_t)(len | TILE_ENTRY_LOOP | (distinct && len <= 8 ? TILE_ENTRY_LOOP8 : 0u)) << 32);
}

constexpr uint64_t TILE_VALID_BIT = 1ull << 51;

// wide = false: the tile kernels' form — BYTE OFFSETS of the three 16-byte rows, pre-shifted so that each is one
//   mask (or shift + mask) away: lo = off0 | off1 << 14, hi = off2 | signs << 24 (off = var·16, var < 2^14)
// wide = true:  three 16-bit variable indices + sign bits at 48..50 (cluster kernel, up to 65 535 variables)
inline uint64_t pack_entry3(const int32_t* var, const bool* neg, bool wide) {
    if (wide) {
        uint64_t e = TILE_VALID_BIT;
        for (int j = 0; j < 3; ++j) {
            e |= (uint64_t)(uint16_t)var[j] << (16 * j);
            if (neg[j]) e |= 1ull << (48 + j);
        }
        return e;
    }
    const uint32_t lo = ((uint32_t)var[0] << 4) | ((uint32_t)var[1] << 18);
    uint32_t hi = (uint32_t)var[2] << 4;
    for (int j = 0; j < 3; ++j) if (neg[j]) hi |= 1u << (24 + j);
    return (uint64_t)lo | ((uint64_t)hi << 32);
}

// Arrange the clauses of one level into slots, 8 per quarter-warp (one 128-byte shared-memory
// wavefront of 16-byte rows), so that the 8 lanes hit 8 distinct bank groups (row index mod 8) in
// each literal position wherever possible.  Literal positions inside a clause may be permuted
// (the arithmetic is symmetric in the literals when the variables are distinct).
// Best-fit over ALL octets of the level: each clause goes to the first octet — and literal
// permutation — where it adds no conflict; clauses that fit nowhere are placed afterwards where
// they cost least.  Returns through `wavefront_sum / wavefront_cnt` the average wavefronts per
// quarter-warp access.
inline void pack_level(const odesat_formula& f, const std::vector<int32_t>& clauses, std::vector<int32_t>& out_perm,
                       std::vector<uint64_t>& out_entry, double& wavefront_sum, int64_t& wavefront_cnt, bool wide) {
    static const int P[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
    const size_t n = clauses.size();
    // Compact: no holes inside a level (measured on B200: spare hole slots improve the packing
    // from 1.23 to 1.12 wavefronts per access but cost more in extra warps and traffic).
    const size_t nb = (n + 7) / 8;
    struct Cl { int32_t m; int32_t var[3]; bool neg[3]; int len; };   // len < 3: positions len.. do not exist (var = -1)
    struct Bin { uint8_t mask[3] = {0, 0, 0}; uint8_t cnt[3][8] = {}; int filled = 0; Cl slot[8]; int cap = 8; };
    std::vector<Bin> bins(nb);
    if (n % 8) bins[nb - 1].cap = (int)(n % 8);   // the level's last octet is partial
    auto load = [&](int32_t m) {
        Cl c;
        c.m = m;
        c.len = (int)std::min<int64_t>(3, f.h_off[m + 1] - f.h_off[m]);
        for (int j = 0; j < 3; ++j) {
            if (j >= c.len) { c.var[j] = -1; c.neg[j] = false; continue; }
            const int32_t l = f.h_lits[f.h_off[m] + j];
            c.var[j] = (l < 0 ? -l : l) - 1;
            c.neg[j] = l < 0;
        }
        return c;
    };
    // a permutation may only move literals that exist (the missing positions stay at the end)
    auto allowed = [&](const Cl& c, int p) {
        for (int j = c.len; j < 3; ++j) if (P[p][j] != j) return false;
        return true;
    };
    auto place = [&](Bin& b, const Cl& c, int p) {
        Cl q;
        q.m = c.m;
        q.len = c.len;
        for (int j = 0; j < 3; ++j) { q.var[j] = c.var[P[p][j]]; q.neg[j] = c.neg[P[p][j]]; }
        for (int j = 0; j < q.len; ++j) { b.mask[j] |= (uint8_t)(1u << (q.var[j] & 7)); b.cnt[j][q.var[j] & 7]++; }
        b.slot[b.filled++] = q;
    };
    std::vector<Cl> leftover;
    size_t first_open = 0;
    for (size_t i = 0; i < n; ++i) {
        const Cl c = load(clauses[i]);
        bool done = false;
        while (first_open < nb && bins[first_open].filled >= bins[first_open].cap) ++first_open;
        for (size_t bi = first_open; bi < nb && !done; ++bi) {
            Bin& b = bins[bi];
            if (b.filled >= b.cap) continue;
            for (int p = 0; p < 6 && !done; ++p) {
                if (!allowed(c, p)) continue;
                bool free = true;
                for (int j = 0; j < c.len; ++j) free = free && !((b.mask[j] >> (c.var[P[p][j]] & 7)) & 1);
                if (free) {
                    place(b, c, p);
                    done = true;
                }
            }
        }
        if (!done) leftover.push_back(c);
    }
    for (const Cl& c : leftover) {   // cheapest (octet, permutation) by added wavefronts
        int best_cost = 1 << 30, best_p = 0;
        size_t best_b = 0;
        for (size_t bi = 0; bi < nb; ++bi) {
            Bin& b = bins[bi];
            if (b.filled >= b.cap) continue;
            for (int p = 0; p < 6; ++p) {
                if (!allowed(c, p)) continue;
                int cost = 0;
                for (int j = 0; j < c.len; ++j) {
                    const int r = c.var[P[p][j]] & 7;
                    int mx = 0;
                    for (int k = 0; k < 8; ++k) mx = std::max<int>(mx, b.cnt[j][k]);
                    if (b.cnt[j][r] + 1 > mx) cost += 1;
                }
                if (cost < best_cost) { best_cost = cost; best_b = bi; best_p = p; }
            }
        }
        place(bins[best_b], c, best_p);
    }
    size_t last_used = 0;
    for (size_t bi = 0; bi < nb; ++bi) if (bins[bi].filled) last_used = bi;
    for (size_t bi = 0; bi <= last_used && n > 0; ++bi) {
        const Bin& b = bins[bi];
        for (int k = 0; k < 8; ++k) {
            if (k < b.filled) {
                out_perm.push_back(b.slot[k].m);
                const Cl& c = b.slot[k];
                if (c.len == 3) out_entry.push_back(pack_entry3(c.var, c.neg, wide));
                else {   // one or two literals (never for the cluster kernel: ragged formulas are not `wide`)
                    int32_t var[3] = {c.var[0], c.len > 1 ? c.var[1] : 0, 0};
                    out_entry.push_back(pack_entry3(var, c.neg, false) |
                                        ((uint64_t)(TILE_ENTRY_NO2 | (c.len < 2 ? TILE_ENTRY_NO1 : 0u)) << 32));
                }
            } else if (bi != last_used) {   // cannot happen with compact bins; kept as a guard
                out_perm.push_back(-1);
                out_entry.push_back(0);
            }
        }
        if (b.filled) {
            for (int j = 0; j < 3; ++j) {
                int mx = 0;
                for (int k = 0; k < 8; ++k) mx = std::max<int>(mx, b.cnt[j][k]);
                wavefront_sum += mx;
                ++wavefront_cnt;
            }
        }
    }
}

// Slots a clause occupies in the schedule: 1, or — a GROUP clause (4..32 literals, distinct variables; one lane per
// literal, tile_ragged.cuh) — its length rounded up to 4 / 8 / 16 / 32.
// Measured on B200 (N = 10 000, 43 000 clauses of which 2 500 have 5 and 500 have 8 literals, 4 096 replicas, f32, ms per
// step; gather engine 1.67): EXACT — one 512-slot item per level, every level waits for its slowest thread — 2.57 with
// the long clauses walked by ONE thread each (clause_loop8: the literal words come from L2, then a chain of eight
// literals), 1.67 as group clauses; BALANCED — wide levels that absorb slow threads — 1.27 with one thread per clause,
// 1.68 as groups (a lane per literal costs as many instructions as a whole packed clause: ncu counts 2.5 x the
// instructions of the uniform kernel).  So: groups under EXACT only.  ODESAT_TILE_GROUPS=0/1 forces them off / on.
inline bool tile_use_groups(int kind) {
    const char* e = std::getenv("ODESAT_TILE_GROUPS");   // read per call: the tests switch it between batches
    return e ? e[0] != '0' : kind == ODESAT_SCHED_EXACT;
}
inline int tile_group_log2(const odesat_formula& f, int64_t m, int kind) {   // 0: not a group clause
    const int64_t len = f.h_off[m + 1] - f.h_off[m];
    if (!tile_use_groups(kind) || len < 4 || len > 32) return 0;
    const int32_t* l = &f.h_lits[f.h_off[m]];
    for (int64_t i = 0; i < len; ++i)
        for (int64_t j = i + 1; j < len; ++j)
            if (std::abs(l[i]) == std::abs(l[j])) return 0;
    return len <= 4 ? 2 : len <= 8 ? 3 : len <= 16 ? 4 : 5;
}

// Level assignment only (shared by every warp count); cached on the formula.
struct TileLevels {
    int nlev = 0;
    std::vector<std::vector<int32_t>> bucket;   // clauses of each level, ascending index
};

// BALANCED: `target` = clauses per level aimed for, `round` = the item width the class capacity is
// rounded up to (a level is a whole number of items wherever possible).
// `ipl` != 1 (BALANCED): wide levels of `ipl` items of `target` clauses each (0 = as many as the degree bound allows).
// EXACT: `target` = cap on the clauses of a level (<= 0: none).
inline std::shared_ptr<TileLevels> build_tile_levels(const odesat_formula& f, int kind, int target = 1024, int round = 512, int ipl = 1) {
    auto s = std::make_shared<TileLevels>();
    const int64_t M = f.M, N = f.N;
    std::vector<int32_t> level(M, 0);
    int nlev = 0;
    auto var_of = [&](int64_t m, int j) {
        const int32_t l = f.h_lits[f.h_off[m] + j];
        return (l < 0 ? -l : l) - 1;
    };
    auto len_of = [&](int64_t m) { return (int)(f.h_off[m + 1] - f.h_off[m]); };   // 3 everywhere for the uniform formulas
    if (kind == ODESAT_SCHED_EXACT) {
        // Order-preserving levels by LIST SCHEDULING: clause m may run once the previous clause of each of
        // its variables has run in an EARLIER level (so every variable still meets its clauses in ascending
        // clause index).  `target` > 0 caps a level at that many clauses — the CTA width, so a level is one
        // item — and the clauses that wait are the ones with the most slack (priority = length of the longest
        // dependency chain hanging off the clause).  On random 3-SAT the level count stays at the critical
        // path (94 at N = 10 000) while the item count drops from 121 to 94 for 512-thread CTAs.
        // target <= 0: as soon as possible, no cap.
        // succ is indexed by literal slot (h_off[m] + j: 3m + j for uniform 3-literal clauses)
        std::vector<int32_t> succ((size_t)f.L, -1), npred(M, 0), last_slot(N, -1), height(M, 0);
        for (int64_t m = 0; m < M; ++m)
            for (int j = 0; j < len_of(m); ++j) {
                const int32_t v = var_of(m, j);
                // (a variable repeated inside a clause is no dependency: a loop clause adds its literals in order)
                if (last_slot[v] >= 0 && last_slot[v] < f.h_off[m]) { succ[last_slot[v]] = (int32_t)m; ++npred[m]; }
                last_slot[v] = (int32_t)(f.h_off[m] + j);
            }
        for (int64_t m = M - 1; m >= 0; --m) {
            int32_t h = 0;
            for (int j = 0; j < len_of(m); ++j) if (succ[f.h_off[m] + j] >= 0) h = std::max(h, height[succ[f.h_off[m] + j]] + 1);
            height[m] = h;
        }
        using Key = std::pair<int32_t, int32_t>;   // (−height, clause): tallest chain first, then lowest index
        std::priority_queue<Key, std::vector<Key>, std::greater<Key>> ready;
        for (int64_t m = 0; m < M; ++m) if (npred[m] == 0) ready.push({-height[m], (int32_t)m});
        std::vector<int32_t> take;
        int64_t placed = 0;
        // the cap counts SLOTS: a group clause takes one per literal (rounded up), everything else one
        std::vector<int32_t> weight(M, 1);
        if (f.K != 3) for (int64_t m = 0; m < M; ++m) { const int lg = tile_group_log2(f, m, kind); if (lg) weight[m] = 1 << lg; }
        while (placed < M) {
            take.clear();
            int64_t slots = 0;
            while (!ready.empty() && (target <= 0 || slots < target)) {
                // (a clause that would overflow the level waits, unless the level is still empty)
                const int32_t m = ready.top().second;
                if (target > 0 && slots > 0 && slots + weight[m] > target) break;
                take.push_back(m);
                slots += weight[m];
                ready.pop();
            }
            for (int32_t m : take) level[m] = nlev;
            for (int32_t m : take)
                for (int j = 0; j < len_of(m); ++j) {
                    const int32_t q = succ[(size_t)f.h_off[m] + j];
                    if (q >= 0 && --npred[q] == 0) ready.push({-height[q], q});
                }
            placed += (int64_t)take.size();
            ++nlev;
        }
    } else {
        // Greedy least-loaded colouring with C colours of capacity `cap`, most constrained clauses first; colours
        // are added when a clause fits nowhere.  → number of colours used.
        std::vector<int32_t> order(M);
        std::iota(order.begin(), order.end(), 0);
        std::vector<int32_t> deg(N);
        for (int64_t i = 0; i < N; ++i) deg[i] = f.h_voff[i + 1] - f.h_voff[i];
        std::vector<int64_t> dsum(M, 0);
        for (int64_t m = 0; m < M; ++m)
            for (int j = 0; j < len_of(m); ++j) dsum[m] += deg[var_of(m, j)];
        std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return dsum[a] > dsum[b]; });
        std::vector<int32_t> weight(M, 1);   // slots per clause (group clauses: one per literal, rounded up)
        int64_t total_slots = M;
        if (f.K != 3) {
            total_slots = 0;
            for (int64_t m = 0; m < M; ++m) { const int lg = tile_group_log2(f, m, kind); if (lg) weight[m] = 1 << lg; total_slots += weight[m]; }
        }
        auto colour = [&](int C, int cap, std::vector<int32_t>& lvl) {
            const int words = (C + 63 + 64) / 64;             // slack for overflow colours
            std::vector<uint64_t> bits((size_t)N * words, 0);
            std::vector<int32_t> load((size_t)words * 64, 0);
            int ncol = C;
            std::vector<uint64_t> forb(words);
            for (int32_t m : order) {
                const int len = len_of(m);
                std::fill(forb.begin(), forb.end(), 0);
                for (int j = 0; j < len; ++j) {
                    const uint64_t* b = &bits[(size_t)var_of(m, j) * words];
                    for (int w = 0; w < words; ++w) forb[w] |= b[w];
                }
                int best = -1, bl = INT32_MAX;
                for (int c = 0; c < ncol; ++c) {
                    if (!((forb[c >> 6] >> (c & 63)) & 1) && load[c] < bl && load[c] < cap) { best = c; bl = load[c]; }
                }
                if (best < 0) {
                    if (ncol >= words * 64) throw Error(ODESAT_EINVAL, "balanced schedule ran out of colours");
                    best = ncol++;
                }
                lvl[m] = best;
                load[best] += weight[m];
                for (int j = 0; j < len; ++j) bits[(size_t)var_of(m, j) * words + (best >> 6)] |= 1ull << (best & 63);
            }
            return ncol;
        };
        if (ipl == 1) {
            // one item per level: levels of `target` clauses, but never fewer colours than the max variable degree
            // (+ 2 of slack for the greedy); classes are capped at a multiple of `round`
            const int64_t S = total_slots;   // = M for uniform formulas
            const int C = (int)std::max<int64_t>(f.max_degree + 2, (S + target - 1) / target);
            const int cap = (int)(((S + C - 1) / C + round - 1) / round * round);
            nlev = colour(C, cap, level);
        } else {
            // WIDE levels (kernels whose ring does not need a barrier after every item): a level is `k` items of
            // `target` clauses, and the colour count goes down to its lower bound, the maximum variable degree —
            // half the level barriers at the headline size (28 levels of 2 × 768 instead of 56 of 768).  k = 0: as
            // many items per level as that bound allows.  The greedy has no slack at C·cap ≈ M, so C, C + 1, C + 2
            // are tried and the one with the fewest items (then levels) is kept.
            const int64_t md = std::max<int64_t>(1, f.max_degree);
            const int64_t M = total_slots;   // (shadows the clause count: the capacities below are in slots)
            // k = 0: the k whose colour count max(md, ceil(M / (k·target))) gives the fewest items (then the fewest levels)
            int64_t k = ipl;
            if (k <= 0) {
                int64_t best_it = INT64_MAX;
                for (int64_t kk = 1; kk <= 4; ++kk) {
                    const int64_t Ck = std::max<int64_t>(md, (M + kk * target - 1) / (kk * target));
                    const int64_t load = (M + Ck - 1) / Ck;
                    const int64_t it = Ck * ((load + target - 1) / target);
                    if (it < best_it || (it == best_it && Ck < std::max<int64_t>(md, (M + k * target - 1) / (k * target)))) { best_it = it; k = kk; }
                }
            }
            const int C0 = (int)std::max<int64_t>(md, (M + k * target - 1) / (k * target));
            int64_t best_items = INT64_MAX;
            int best_lev = 0;
            std::vector<int32_t> lvl(f.M, 0);
            for (int C = C0; C <= C0 + 2; ++C) {
                const int nc = colour(C, (int)(k * target), lvl);
                std::vector<int32_t> cnt(nc, 0);
                for (int64_t m = 0; m < f.M; ++m) cnt[lvl[m]] += weight[m];
                int64_t items = 0;
                int lev = 0;
                for (int c : cnt) { items += (c + target - 1) / target; lev += c > 0; }
                if (items < best_items || (items == best_items && lev < best_lev)) {
                    best_items = items; best_lev = lev; level = lvl; nlev = nc;
                }
            }
        }
    }
    s->nlev = nlev;
    s->bucket.assign(nlev, {});
    for (int64_t m = 0; m < M; ++m) s->bucket[level[m]].push_back((int32_t)m);
    return s;
}

// Items per BALANCED level the tile engine compiles for: 1 = one CTA width of clauses per level (a barrier after every
// item), 0 (default) = wide levels, as few colours as the maximum variable degree allows.  ODESAT_TILE_IPL overrides.
inline int tile_items_per_level() {
    const char* e = std::getenv("ODESAT_TILE_IPL");
    return e ? std::max(0, std::atoi(e)) : 0;
}
// The BALANCED levels of a CTA of `cap` threads, as the tile engine builds them.
inline std::shared_ptr<TileLevels> build_balanced_levels(const odesat_formula& f, int cap, int ipl) {
    const int target = cap >= 512 ? cap : 1024;
    return (cap >= 512 && ipl != 1) ? build_tile_levels(f, ODESAT_SCHED_BALANCED, target, target, ipl)
                                    : build_tile_levels(f, ODESAT_SCHED_BALANCED, target, target / 2);
}

// `depth`: prefetch ring depth of the kernel the schedule is for; the item list is padded with
// empty items to a multiple of it (and to more than one ring) so that item i always lives in
// ring slot i % depth and a slot is stored before it is prefetched again.
inline std::shared_ptr<TileSchedule> build_tile_schedule(const odesat_formula& f, const TileLevels& lv, int kind, int nt, int depth,
                                                         bool upload = true, bool wide = false) {
    auto s = std::make_shared<TileSchedule>();
    if (!wide && f.N >= (1 << 14)) throw Error(ODESAT_EUNSUPPORTED, "tile schedule: more than 16383 variables");
    s->kind = kind;
    s->nt = nt;
    s->M = f.M;
    double wsum = 0;
    int64_t wcnt = 0;
    std::vector<int32_t> three, loop, group;
    for (const auto& b : lv.bucket) {
        if (b.empty()) continue;
        const size_t base0 = s->perm.size();
        if (f.K == 3 && f.distinct_vars) pack_level(f, b, s->perm, s->entry, wsum, wcnt, wide);
        else {
            // ragged formula: the 3-literal clauses (with three distinct variables) are packed as usual; the others
            // follow them (so that they share warps) as LOOP clauses whose literals live in `aux`
            if (wide) throw Error(ODESAT_EUNSUPPORTED, "tile schedule: the cluster kernel needs uniform 3-literal clauses");
            three.clear();
            loop.clear();
            group.clear();
            auto plain3 = [&](int32_t m) {   // one to three literals with distinct variables: the packed word
                const int64_t len = f.h_off[m + 1] - f.h_off[m];
                if (len < 1 || len > 3) return false;
                const int32_t* l = &f.h_lits[f.h_off[m]];
                for (int64_t i = 0; i < len; ++i)
                    for (int64_t j = i + 1; j < len; ++j)
                        if (std::abs(l[i]) == std::abs(l[j])) return false;
                return true;
            };
            for (int32_t m : b) {
                if (plain3(m)) three.push_back(m);
                else if (tile_group_log2(f, m, kind)) group.push_back(m);
                else loop.push_back(m);
            }
            // group clauses first, largest first: packed from the level's (8-slot aligned) start, groups of descending
            // power-of-two size are aligned to their size inside the warps (thread = slot - level start, mod the CTA width)
            auto log2p_of = [&](int32_t m) { return (uint32_t)tile_group_log2(f, m, kind); };
            std::stable_sort(group.begin(), group.end(), [&](int32_t x, int32_t y) { return log2p_of(x) > log2p_of(y); });
            for (int32_t m : group) {
                const uint32_t len = (uint32_t)(f.h_off[m + 1] - f.h_off[m]), lp = log2p_of(m);
                for (uint32_t j = 0; j < (1u << lp); ++j) {
                    s->perm.push_back(j == 0 ? m : -1);
                    if (j < len) {
                        const int32_t l = f.h_lits[f.h_off[m] + j];
                        s->entry.push_back(pack_entry_group((uint32_t)((l < 0 ? -l : l) - 1), l < 0, lp, j, true));
                    } else s->entry.push_back(pack_entry_group(0, false, lp, j, false));
                }
                ++s->n_group;
            }
            // holes up to the next octet (pack_level's unit): void non-leader lanes, which do nothing
            while ((s->perm.size() - base0) % 8) { s->perm.push_back(-1); s->entry.push_back(pack_entry_group(0, false, 0, 1, false)); }
            if (!three.empty()) pack_level(f, three, s->perm, s->entry, wsum, wcnt, wide);
            for (int32_t m : loop) {
                const uint32_t len = (uint32_t)(f.h_off[m + 1] - f.h_off[m]);
                if (len > 0xFFFFu) throw Error(ODESAT_EUNSUPPORTED, "tile schedule: a clause has more than 65535 literals");
                s->perm.push_back(m);
                bool distinct = true;
                for (uint32_t i = 0; i < len && distinct; ++i)
                    for (uint32_t j = i + 1; j < len && distinct; ++j)
                        distinct = std::abs(f.h_lits[f.h_off[m] + i]) != std::abs(f.h_lits[f.h_off[m] + j]);
                s->entry.push_back(pack_entry_loop((uint32_t)s->aux.size(), len, distinct));   // offset: a multiple of 4 words
                for (uint32_t j = 0; j < len; ++j) {
                    const int32_t l = f.h_lits[f.h_off[m] + j];
                    s->aux.push_back(((uint32_t)((l < 0 ? -l : l) - 1) << 4) | (l < 0 ? 0x80000000u : 0u));
                }
                if (len == 0) s->aux.push_back(0u);                                    // the kernel reads 16-byte vectors,
                while (s->aux.size() % 4) s->aux.push_back(0u);                        // the first one unconditionally
                ++s->n_loop;
            }
        }
        const size_t n = s->perm.size() - base0;
        while (s->perm.size() % 8) { s->perm.push_back(-1); s->entry.push_back(0); }   // 128-byte aligned level start
        for (size_t o = 0; o < n; o += (size_t)nt) {
            const size_t cnt = std::min<size_t>((size_t)nt, n - o);
            const bool last = o + (size_t)nt >= n;
            s->items.push_back(pack_item((uint32_t)((base0 + o) & 0xFFFFFu), (uint32_t)(cnt & 0x7FFu), last));
            s->items2.push_back(make_uint2((uint32_t)(base0 + o), (uint32_t)cnt | (last ? TILE_ITEM_LAST : 0u)));
        }
        ++s->nlev;
    }
    const size_t pad = (size_t)((depth % 2) ? 2 * depth : depth);   // the strict first-step kernel uses a ring of 2
    while (s->items.size() % pad || s->items.size() <= (size_t)depth) {
        s->items.push_back(pack_item(0, 0, false));
        s->items2.push_back(make_uint2(0u, 0u));
    }
    s->Mpad = (int64_t)s->perm.size();
    if (nt <= 1024 && s->Mpad >= (1 << 20)) throw Error(ODESAT_EUNSUPPORTED, "tile schedule: more than 2^20 clause slots");
    s->n_items = (int)s->items.size();
    s->conflict_wavefronts = wcnt ? wsum / wcnt : 1.0;
    if (!upload) return s;
    s->d_items.alloc(std::max<size_t>(s->items.size(), 1));
    s->d_items2.alloc(std::max<size_t>(s->items2.size(), 1));
    s->d_perm.alloc(std::max<size_t>(s->perm.size(), 1));
    s->d_entry.alloc(std::max<size_t>(s->entry.size(), 1));
    s->d_aux.alloc(std::max<size_t>(s->aux.size(), 1));
    if (!s->aux.empty()) ODESAT_CUDA(cudaMemcpy(s->d_aux.p, s->aux.data(), s->aux.size() * 4, cudaMemcpyHostToDevice));
    if (!s->perm.empty()) {
        ODESAT_CUDA(cudaMemcpy(s->d_items.p, s->items.data(), s->items.size() * 4, cudaMemcpyHostToDevice));
        ODESAT_CUDA(cudaMemcpy(s->d_items2.p, s->items2.data(), s->items2.size() * 8, cudaMemcpyHostToDevice));
        ODESAT_CUDA(cudaMemcpy(s->d_perm.p, s->perm.data(), s->perm.size() * 4, cudaMemcpyHostToDevice));
        ODESAT_CUDA(cudaMemcpy(s->d_entry.p, s->entry.data(), s->entry.size() * 8, cudaMemcpyHostToDevice));
    }
    return s;
}

}  // namespace odesat
