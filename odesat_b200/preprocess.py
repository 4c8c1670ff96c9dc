"""Ratio preprocessing of `solve`: bounded variable elimination + blocked
clause elimination + subsumption until the clause/variable ratio reaches a target, and the
trace replay that re-derives the eliminated variables afterwards.

Host-side restatement of cnf.rs:317-840 (sequential set algebra, not a GPU path).  Clauses are
frozensets of signed DIMACS literals; wherever the reference iterates a BTreeSet the order is
reproduced (literals ordered by (variable, is_negated), clauses lexicographically); wherever it
iterates a HashSet/HashMap (cnf.rs:728, 780 — arbitrary order, so the reference's own output
differs from run to run) ascending variable order is used.  Only "a valid elimination sequence
reaching the ratio" can therefore be matched, not one particular run of the reference.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, FrozenSet, Iterable, List, Optional, Set, Tuple

import numpy as np

Clause = FrozenSet[int]


def _lit_key(l: int) -> Tuple[int, bool]:
    return (abs(l), l < 0)          # derive(Ord) on Literal { variable, is_negated } (cnf.rs:5-9)


def _clause_key(c: Clause):
    return sorted(_lit_key(l) for l in c)


def sorted_clauses(clauses: Iterable[Clause]) -> List[Clause]:
    """BTreeSet<CNFClauseSet> iteration order."""
    return sorted(clauses, key=_clause_key)


def sorted_literals(c: Clause) -> List[int]:
    return sorted(c, key=_lit_key)


@dataclass
class Trace:
    """cnf.rs:558-585: ('ve', var, modified positive clauses) | ('bce', var, clause)."""
    steps: List[tuple] = field(default_factory=list)


Index = Dict[int, Tuple[Set[Clause], Set[Clause]]]


def calculate_variable_indices(clauses: Iterable[Clause]) -> Index:
    """cnf.rs:418-438: variable → (clauses with the positive literal, with the negative literal)."""
    idx: Index = {}
    for c in clauses:
        for l in c:
            pos, neg = idx.setdefault(abs(l), (set(), set()))
            (neg if l < 0 else pos).add(c)
    return idx


def is_tautology(c: Clause) -> bool:
    """cnf.rs:541-551."""
    return any(-l in c for l in c)


def calculate_resolvents(idx: Index, clause: Clause, variable: int) -> List[Clause]:
    """cnf.rs:440-479: resolvents of `clause` on `variable`; tautological (w.r.t. `clause`) and
    empty resolvents are dropped, exactly as the reference does.  (The reference walks both sets in
    BTreeSet order; every caller uses the result as a set / under `all`, so the order is immaterial.)"""
    others = idx[variable][1] if variable in clause else idx[variable][0]
    base = frozenset(l for l in clause if abs(l) != variable)
    neg_base = frozenset(-l for l in base)
    out: List[Clause] = []
    for other in others:
        rest = [l for l in other if abs(l) != variable]
        if not neg_base.isdisjoint(rest):          # a literal of `other` contradicts one of `clause`: cleared
            continue
        combined = base.union(rest)
        if combined:
            out.append(combined)
    return out


def calculate_var_resolvents(idx: Index, variable: int) -> Set[Clause]:
    """cnf.rs:481-498."""
    out: Set[Clause] = set()
    for pc in idx[variable][0]:
        out.update(calculate_resolvents(idx, pc, variable))
    return out


def subsume_clauses(clauses: Set[Clause]) -> None:
    """cnf.rs:521-539: drop every clause that is a proper superset of another one.  The result does
    not depend on the visiting order.  A proper superset is strictly longer, so clauses are visited by
    increasing length and each one is tested only against the shorter clauses filed under one of ITS
    literals (a subset's filing literal must occur in the superset), as bit masks."""
    bit: Dict[int, int] = {}
    by_len: Dict[int, List[Tuple[int, Clause]]] = {}
    for c in clauses:
        m = 0
        for l in c:
            b = bit.get(l)
            if b is None:
                b = bit[l] = 1 << len(bit)
            m |= b
        by_len.setdefault(len(c), []).append((m, c))
    filed: Dict[int, List[int]] = {}           # literal → masks of the shorter clauses filed under it
    drop: List[Clause] = []
    for n in sorted(by_len):
        for m, c in by_len[n]:
            hit = False
            for l in c:
                for p in filed.get(l, ()):
                    if p & m == p:
                        hit = True
                        break
                if hit:
                    break
            if hit:
                drop.append(c)
        for m, c in by_len[n]:
            if c:
                filed.setdefault(min(c), []).append(m)
            else:                               # the empty clause is a subset of every clause
                for l in bit:
                    filed.setdefault(l, []).append(m)
    for c in drop:
        clauses.discard(c)


def is_blocked(clause: Clause, idx: Index) -> Optional[int]:
    """cnf.rs:588-599."""
    for l in sorted_literals(clause):
        if all(is_tautology(r) for r in calculate_resolvents(idx, clause, abs(l))):
            return abs(l)
    return None


def _eliminate_if_blocked(clause: Clause, clauses: Set[Clause], idx: Index):
    """cnf.rs:602-631."""
    var = is_blocked(clause, idx)
    if var is None:
        return None
    changed = set()
    for l in clause:
        changed.add(abs(l))
        pos, neg = idx.setdefault(abs(l), (set(), set()))
        (neg if l < 0 else pos).discard(clause)
    clauses.discard(clause)
    return changed, ("bce", var, clause)


def _eliminate_variable(clauses: Set[Clause], idx: Index, variable: int, resolvents: Set[Clause]):
    """cnf.rs:634-715 → (changed variables, positive clauses with the literal removed)."""
    if variable not in idx:
        return set(), set()
    pos, neg = idx.pop(variable)
    originals = pos | neg
    changed = {abs(l) for c in originals for l in c}
    for v in changed:
        if v in idx:
            p, n = idx[v]
            p.difference_update(originals)
            n.difference_update(originals)
    clauses.difference_update(originals)
    clauses.update(resolvents)
    for r in resolvents:
        for l in r:
            p, n = idx.setdefault(abs(l), (set(), set()))
            (n if l < 0 else p).add(r)
    return changed, {frozenset(c - {variable}) for c in pos}


def _min_ratio_resolvant(variables: Set[int], idx: Index, n_clauses: int, varnum: int, target: np.float32):
    """cnf.rs:718-754 (f32 ratio arithmetic; ties: the first variable in ascending order)."""
    best, smallest = None, np.float32(np.finfo(np.float32).max)
    for v in sorted(variables):
        if v not in idx:
            continue
        pos, neg = idx[v]
        res = {r for r in calculate_var_resolvents(idx, v) if not is_tautology(r)}
        subsume_clauses(res)
        count = n_clauses - len(pos) - len(neg) + len(res)
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = np.float32(count) / np.float32(varnum - 1)
        if ratio < smallest:
            smallest, best = ratio, (v, res)
    return None if smallest > target else best


def repeatedly_resolve_and_update(clauses: Iterable[Clause], varnum: int, desired_ratio: float, log=None):
    """cnf.rs:833-840 + 756-829 → (clauses, varnum, trace)."""
    clauses = set(clauses)
    idx = calculate_variable_indices(clauses)
    trace = Trace()
    target = np.float32(desired_ratio)
    for c in [c for c in sorted_clauses(clauses) if is_blocked(c, idx) is not None]:
        r = _eliminate_if_blocked(c, clauses, idx)
        if r:
            trace.steps.append(r[1])
    elim = set(idx)
    while True:
        pick = _min_ratio_resolvant(elim, idx, len(clauses), varnum, target)
        if pick is None:
            break
        variable, res = pick
        elim, modified = _eliminate_variable(clauses, idx, variable, res)
        varnum -= 1                                              # cnf.rs:685
        trace.steps.append(("ve", variable, modified))
        for r in sorted_clauses(res):
            b = _eliminate_if_blocked(r, clauses, idx)
            if b:
                trace.steps.append(b[1])
                elim |= b[0]
    subsume_clauses(clauses)
    if log:
        log(f"Clauses: {len(clauses)} | Vars: {varnum}")          # cnf.rs:822-826
    return clauses, varnum, trace


def _evaluate_cnf_set(assign: Dict[int, bool], clauses: Iterable[Clause]) -> bool:
    """cnf.rs:266-287 — missing variables are INSERTED as false, like `entry().or_insert(false)`."""
    for c in sorted_clauses(clauses):
        ok = False
        for l in sorted_literals(c):
            val = assign.setdefault(abs(l), False)
            ok = ok or (not val if l < 0 else val)
        if not ok:
            return False
    return True


def calculate_trace(assign: Dict[int, bool], trace: Trace) -> None:
    """cnf.rs:501-519: replay the trace backwards, re-deriving eliminated variables in place."""
    for step in reversed(trace.steps):
        if step[0] == "ve":
            _, var, cls = step
            assign[var] = not _evaluate_cnf_set(assign, cls)
        else:
            _, var, clause = step
            if not _evaluate_cnf_set(assign, [clause]):
                assign[var] = not assign.get(var, False)


def to_clause_set(raw_clauses: Iterable[Iterable[int]]) -> Set[Clause]:
    """cnf.rs:381-394: the BTreeSet conversion deduplicates clauses and literals."""
    return {frozenset(c) for c in raw_clauses}
