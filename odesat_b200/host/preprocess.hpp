// preprocess.hpp — the ratio preprocessing of `solve` (cnf.rs:317-840) restated on the host in C++:
// bounded variable elimination + blocked clause elimination + subsumption until the clause / variable
// ratio reaches the target, and the trace replay that re-derives the eliminated variables.
//
// Sequential set algebra, not a GPU path.  A literal is the code 2·var + negated,
// whose integer order is the derive(Ord) order of `Literal { variable, is_negated }` (cnf.rs:5-9); a
// clause is the sorted vector of its codes (a BTreeSet<Literal>), and clauses compare
// lexicographically (BTreeSet<CNFClauseSet>).  Wherever the reference iterates a HashSet / HashMap
// (cnf.rs:728, 780 — arbitrary order, so its own output differs from run to run) ascending variable
// order is used, the same choice as odesat_b200/preprocess.py, which this file matches result for result.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <limits>
#include <map>
#include <set>
#include <utility>
#include <vector>

namespace odesat {
namespace prep {

using Lit = uint32_t;                       // 2·var + (negated ? 1 : 0)
using Clause = std::vector<Lit>;            // sorted, unique
using ClauseSet = std::set<Clause>;
using Index = std::map<std::size_t, std::pair<ClauseSet, ClauseSet>>;   // var → (with +var, with −var)

inline Lit make_lit(std::size_t var, bool neg) { return (Lit)(2 * var + (neg ? 1 : 0)); }
inline std::size_t var_of(Lit l) { return l >> 1; }
inline bool neg_of(Lit l) { return l & 1u; }
inline bool has(const Clause& c, Lit l) { return std::binary_search(c.begin(), c.end(), l); }

inline Clause make_clause(const std::vector<int>& dimacs) {   // cnf.rs:381-394: duplicates collapse
    Clause c;
    for (int l : dimacs) c.push_back(make_lit((std::size_t)(l < 0 ? -l : l), l < 0));
    std::sort(c.begin(), c.end());
    c.erase(std::unique(c.begin(), c.end()), c.end());
    return c;
}

struct Step {                               // cnf.rs:558-562
    bool ve;                                // VariableElimination(var, clauses) | BlockedClauseElimination(var, clause)
    std::size_t var;
    ClauseSet clauses;
};
using Trace = std::vector<Step>;

inline Index calculate_variable_indices(const ClauseSet& clauses) {   // cnf.rs:418-438
    Index idx;
    for (const Clause& c : clauses)
        for (Lit l : c) (neg_of(l) ? idx[var_of(l)].second : idx[var_of(l)].first).insert(c);
    return idx;
}

inline bool is_tautology(const Clause& c) {   // cnf.rs:541-551
    for (Lit l : c) if (has(c, l ^ 1u)) return true;
    return false;
}

// cnf.rs:440-479: resolvents of `clause` on `variable`; a resolvent that would be tautological and the
// empty resolvent are dropped, exactly as the reference does.
inline std::vector<Clause> calculate_resolvents(const Index& idx, const Clause& clause, std::size_t variable) {
    std::vector<Clause> out;
    const auto& entry = idx.at(variable);
    const ClauseSet& others = has(clause, make_lit(variable, false)) ? entry.second : entry.first;
    Clause base;
    for (Lit l : clause) if (var_of(l) != variable) base.push_back(l);
    for (const Clause& other : others) {
        Clause combined = base;
        bool cleared = false;
        for (Lit l : other) {
            if (var_of(l) == variable) continue;
            if (has(base, l ^ 1u)) { cleared = true; break; }
            combined.push_back(l);
        }
        if (cleared) continue;
        std::sort(combined.begin(), combined.end());
        combined.erase(std::unique(combined.begin(), combined.end()), combined.end());
        if (!combined.empty()) out.push_back(combined);
    }
    return out;
}

inline ClauseSet calculate_var_resolvents(const Index& idx, std::size_t variable) {   // cnf.rs:481-498
    ClauseSet all;
    for (const Clause& pc : idx.at(variable).first)
        for (Clause& r : calculate_resolvents(idx, pc, variable)) all.insert(std::move(r));
    return all;
}

// cnf.rs:521-539: drop every clause that is a proper superset of another one.  The result does not depend on the
// visiting order.  A proper superset is strictly longer, so clauses are visited by increasing length and each is
// tested only against the shorter clauses filed under one of ITS literals (a subset's filing literal — its
// smallest — must occur in the superset), with a 64-bit signature as a first filter.
inline void subsume_clauses(ClauseSet& clauses) {
    struct Item { const Clause* c; uint64_t sig; };
    auto sig_of = [](const Clause& c) { uint64_t s = 0; for (Lit l : c) s |= uint64_t(1) << (l & 63u); return s; };
    std::map<std::size_t, std::vector<Item>> by_len;
    for (const Clause& c : clauses) by_len[c.size()].push_back(Item{&c, sig_of(c)});
    std::map<Lit, std::vector<Item>> filed;
    std::vector<Item> empties;                      // the empty clause is a subset of every clause
    std::vector<const Clause*> drop;
    for (auto& kv : by_len) {
        for (const Item& it : kv.second) {
            bool hit = !empties.empty() && !it.c->empty();
            for (std::size_t k = 0; k < it.c->size() && !hit; ++k) {
                auto f = filed.find((*it.c)[k]);
                if (f == filed.end()) continue;
                for (const Item& p : f->second)
                    if ((p.sig & ~it.sig) == 0 && std::includes(it.c->begin(), it.c->end(), p.c->begin(), p.c->end())) { hit = true; break; }
            }
            if (hit) drop.push_back(it.c);
        }
        for (const Item& it : kv.second) {
            if (it.c->empty()) empties.push_back(it);
            else filed[it.c->front()].push_back(it);
        }
    }
    std::vector<Clause> gone;
    for (const Clause* c : drop) gone.push_back(*c);
    for (const Clause& c : gone) clauses.erase(c);
}

inline bool is_blocked(const Clause& clause, const Index& idx, std::size_t* var) {   // cnf.rs:588-599
    for (Lit l : clause) {
        bool all_taut = true;
        for (const Clause& r : calculate_resolvents(idx, clause, var_of(l))) all_taut = all_taut && is_tautology(r);
        if (all_taut) { *var = var_of(l); return true; }
    }
    return false;
}

// cnf.rs:602-631 → true when the clause was blocked and removed
inline bool eliminate_if_blocked(const Clause& clause, ClauseSet& clauses, Index& idx, std::set<std::size_t>* changed, Step* step) {
    std::size_t var = 0;
    if (!is_blocked(clause, idx, &var)) return false;
    for (Lit l : clause) {
        if (changed) changed->insert(var_of(l));
        auto& e = idx[var_of(l)];
        (neg_of(l) ? e.second : e.first).erase(clause);
    }
    clauses.erase(clause);
    step->ve = false;
    step->var = var;
    step->clauses = ClauseSet{clause};
    return true;
}

// cnf.rs:634-715 → (changed variables, positive clauses with the literal removed)
inline void eliminate_variable(ClauseSet& clauses, Index& idx, std::size_t variable, const ClauseSet& resolvents,
                               std::set<std::size_t>& changed, ClauseSet& modified_pos) {
    changed.clear();
    modified_pos.clear();
    auto it = idx.find(variable);
    if (it == idx.end()) return;
    const ClauseSet pos = it->second.first, neg = it->second.second;
    idx.erase(it);
    for (const ClauseSet* s : {&pos, &neg})
        for (const Clause& c : *s)
            for (Lit l : c) changed.insert(var_of(l));
    for (std::size_t v : changed) {
        auto e = idx.find(v);
        if (e == idx.end()) continue;
        for (const ClauseSet* s : {&pos, &neg})
            for (const Clause& c : *s) { e->second.first.erase(c); e->second.second.erase(c); }
    }
    for (const Clause& c : pos) clauses.erase(c);
    for (const Clause& c : neg) clauses.erase(c);
    for (const Clause& r : resolvents) {
        clauses.insert(r);
        for (Lit l : r) (neg_of(l) ? idx[var_of(l)].second : idx[var_of(l)].first).insert(r);
    }
    for (const Clause& c : pos) {
        Clause m;
        for (Lit l : c) if (l != make_lit(variable, false)) m.push_back(l);
        modified_pos.insert(m);
    }
}

// cnf.rs:718-754 (f32 ratio arithmetic; ties: the first variable in ascending order)
inline bool min_ratio_resolvant(const std::set<std::size_t>& variables, const Index& idx, std::size_t n_clauses,
                                std::size_t varnum, float target, std::size_t* best_var, ClauseSet* best_res) {
    float smallest = std::numeric_limits<float>::max();
    bool found = false;
    for (std::size_t v : variables) {
        auto e = idx.find(v);
        if (e == idx.end()) continue;
        ClauseSet res = calculate_var_resolvents(idx, v);
        for (auto it = res.begin(); it != res.end();) it = is_tautology(*it) ? res.erase(it) : std::next(it);
        subsume_clauses(res);
        const std::size_t count = n_clauses - e->second.first.size() - e->second.second.size() + res.size();
        const float ratio = (float)count / (float)(varnum - 1);
        if (ratio < smallest) {
            smallest = ratio;
            *best_var = v;
            *best_res = std::move(res);
            found = true;
        }
    }
    return found && !(smallest > target);
}

// cnf.rs:833-840 + 756-829: clauses and varnum are reduced in place; prints the reference's progress line.
inline Trace repeatedly_resolve_and_update(ClauseSet& clauses, std::size_t& varnum, float desired_ratio, bool log = true) {
    Index idx = calculate_variable_indices(clauses);
    Trace trace;
    std::vector<Clause> blocked;
    std::size_t dummy = 0;
    for (const Clause& c : clauses) if (is_blocked(c, idx, &dummy)) blocked.push_back(c);
    for (const Clause& c : blocked) {
        Step s;
        if (eliminate_if_blocked(c, clauses, idx, nullptr, &s)) trace.push_back(std::move(s));
    }
    std::set<std::size_t> elim;
    for (const auto& kv : idx) elim.insert(kv.first);
    for (;;) {
        std::size_t variable = 0;
        ClauseSet res;
        if (!min_ratio_resolvant(elim, idx, clauses.size(), varnum, desired_ratio, &variable, &res)) break;
        ClauseSet modified;
        eliminate_variable(clauses, idx, variable, res, elim, modified);
        --varnum;                                                   // cnf.rs:685
        trace.push_back(Step{true, variable, std::move(modified)});
        for (const Clause& r : res) {
            Step s;
            std::set<std::size_t> ch;
            if (eliminate_if_blocked(r, clauses, idx, &ch, &s)) {
                trace.push_back(std::move(s));
                elim.insert(ch.begin(), ch.end());
            }
        }
    }
    subsume_clauses(clauses);
    if (log) std::printf("Clauses: %zu | Vars: %zu\n", clauses.size(), varnum);   // cnf.rs:822-826
    return trace;
}

// cnf.rs:266-287 — missing variables are INSERTED as false, like `entry().or_insert(false)`
inline bool evaluate_cnf_set(std::map<std::size_t, bool>& assign, const ClauseSet& clauses) {
    for (const Clause& c : clauses) {
        bool ok = false;
        for (Lit l : c) {
            const bool val = assign.emplace(var_of(l), false).first->second;
            ok = ok || (neg_of(l) ? !val : val);
        }
        if (!ok) return false;
    }
    return true;
}

// cnf.rs:501-519: replay the trace backwards, re-deriving eliminated variables in place
inline void calculate_trace(std::map<std::size_t, bool>& assign, const Trace& trace) {
    for (auto it = trace.rbegin(); it != trace.rend(); ++it) {
        if (it->ve) {
            assign[it->var] = !evaluate_cnf_set(assign, it->clauses);
        } else if (!evaluate_cnf_set(assign, it->clauses)) {
            assign[it->var] = !assign[it->var];
        }
    }
}

}  // namespace prep
}  // namespace odesat
