"""Host plumbing: DIMACS reader / normaliser semantics of cnf.rs:138-219 and the generator."""
import numpy as np
import pytest

from odesat_b200 import cnf


def test_parser_line_rules():
    text = "c comment\nc\np cnf 5 3\n1 -5 4 0\n-1 5 3 4 0\n-3 -4 0"          # no trailing newline
    f = cnf.parse_dimacs_format(text)
    assert f.varnum == 5 and f.clauses == [[1, -5, 4], [-1, 5, 3, 4], [-3, -4]]
    # blank lines become empty clauses; tokens after the first 0 are dropped; \r\n tolerated
    g = cnf.parse_dimacs_format("p cnf 3 2\r\n1 2 0 3\r\n\r\n-3 0\n")
    assert g.clauses == [[1, 2], [], [-3]]
    # no header → varnum = distinct variables (CNFFormula::new(.., None), cnf.rs:60-76)
    h = cnf.parse_dimacs_format("7 -9 0\n9 0\n")
    assert h.varnum == 2
    with pytest.raises(ValueError):
        cnf.parse_dimacs_format("p cnf 1 1\n%\n")                            # the reference's unwrap() panics


def test_normaliser_keeps_header_varnum():
    f = cnf.normalize_cnf_variables(cnf.parse_dimacs_format("p cnf 10 2\n4 -9 0\n9 2 0\n"))
    assert f.varnum == 10                                                     # cnf.rs:198
    assert f.name_map == {2: 0, 4: 1, 9: 2}
    assert list(f.lits) == [2, -3, 3, 1] and list(f.clause_off) == [0, 2, 4]
    with pytest.raises(ValueError):
        cnf.normalize_cnf_variables(cnf.parse_dimacs_format("p cnf 1 1\n1 2 0\n"))


def test_evaluate_and_render():
    f = cnf.normalize_cnf_variables(cnf.parse_dimacs_format("p cnf 3 3\n1 -2 0\n2 3 0\n-1 -3 0\n"))
    assert f.evaluate([1, 1, 0]) and not f.evaluate([1, 0, 0])
    assert cnf.render_variable_map(f.map_values_by_indices([1, 1, 0])) == "1 1\n2 1\n3 0\n"
    e = cnf.normalize_cnf_variables(cnf.parse_dimacs_format("p cnf 1 2\n1 0\n\n"))
    assert not e.evaluate([1])                                                # an empty clause is false


def test_random_ksat_shape_and_determinism():
    f = cnf.random_ksat(1000, 4.3, seed=20240611)
    assert f.varnum == 1000 and f.n_clauses == 4300 and f.n_literals == 12900
    var = (np.abs(f.lits) - 1).reshape(-1, 3)
    s = np.sort(var, axis=1)
    assert (s[:, 1:] != s[:, :-1]).all()                                      # distinct inside a clause
    assert var.min() >= 0 and var.max() < 1000
    g = cnf.random_ksat(1000, 4.3, seed=20240611)
    assert np.array_equal(f.lits, g.lits)
    assert abs((f.lits < 0).mean() - 0.5) < 0.03
    assert cnf.parse_dimacs_format(cnf.to_dimacs(f)).clauses[0] == [int(x) for x in f.lits[:3]]
