"""The reference's command bodies (main.rs:143-386) end to end on the GPU, including `solve`'s
ratio preprocessing + trace replay (BASELINE.json configs[0]: easy.cnf via `solve -r 7`)."""
import pytest

from odesat_b200 import _lib as L
from odesat_b200 import cnf, commands

pytestmark = pytest.mark.gpu


def _check_render(res, original_path):
    f = cnf.parse_dimacs_format(open(original_path).read())
    values = {int(a): int(b) for a, b in (l.split() for l in res.rendered.splitlines())}
    assert all(any((values.get(abs(l), 0) == 1) != (l < 0) for l in c) for c in f.clauses)


@pytest.mark.parametrize("ratio", [None, 3.0])
def test_solve_with_ratio_preprocessing_config0(golden_dir, tmp_path, ratio):
    lines = []
    out = tmp_path / "a.txt"
    res = commands.solve(str(golden_dir / "aim100_sat.cnf"), output=str(out), ctv_ratio=ratio, step_number=200000,
                         seed=5, log=lines.append)
    assert res.is_satisfiable
    assert res.n_vars < 100                                   # variables were eliminated and re-derived by the trace
    want = ["Reading CNF formula from file...", "Parsing CNF formula...", "Preprocessing CNF formula...",
            f"Clauses: {res.n_clauses} | Vars: {res.n_vars}", "Simulating...", "Mapping values...", "Evaluating CNF formula...",
            "Checking if solution vector satisfies formula: true", "Rendering variable assignments...",
            "Writing results to file..."]
    assert lines == want
    assert out.read_text() == res.rendered
    _check_render(res, golden_dir / "aim100_sat.cnf")


def test_batch_and_inter_commands(golden_dir):
    res = commands.batch(str(golden_dir / "aim100_unsat.cnf"), 1000, 100, step_size=0.01, precision=L.F32, log=lambda s: None)
    assert not res.is_satisfiable and res.steps == 1000 and res.winner == -1      # configs[1]
    res = commands.inter(str(golden_dir / "aim100_sat.cnf"), 128, step_number=6000, step_size=0.01, seed=2, log=lambda s: None)
    assert res.is_satisfiable and res.winner >= 0
    _check_render(res, golden_dir / "aim100_sat.cnf")
    # without -s `inter` is adaptive with ONE dt shared by the replicas (quirk Q7) — offered, sequential like the reference
    res = commands.inter(str(golden_dir / "aim100_sat.cnf"), 4, step_number=20000, seed=3, log=lambda s: None)
    assert res.is_satisfiable and res.winner >= 0


def test_inter_driver_script_single_process():
    """scripts/inter_multi_gpu.py (the torchrun program of SURVEY §8e) with one rank: early exit, winner
    verified on device and host, and the same winner as the single-batch reference run (--check)."""
    import json
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, str(root / "scripts" / "inter_multi_gpu.py"), "--vars", "2000", "--alpha", "3.0",
                        "--dt", "0.04", "--replicas", "64", "--max-steps", "6000", "--check"],
                       capture_output=True, text=True, timeout=300, cwd=str(root))
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["winner"] >= 0 and d["winner_verified_sat"] is True and d["single_gpu_same_winner"] is True
    assert d["flag_step"] < d["steps_run"] <= d["flag_step"] + d["chunk"]


@pytest.mark.parametrize("ratio", [None, 3.0])
def test_stoch_command(golden_dir, tmp_path, ratio):
    """main.rs:206-252: ratio preprocessing → src/stoch.rs search on the GPU → trace replay → exact verification."""
    lines = []
    res = commands.stoch(str(golden_dir / "aim100_sat.cnf"), step_number=200000, ctv_ratio=ratio, batch_size=64, seed=4,
                         log=lines.append)
    assert res.is_satisfiable and res.winner >= 0 and res.n_vars < 100
    assert lines[:3] == ["Reading CNF formula from file...", "Parsing CNF formula...", "Preprocessing CNF formula..."]
    assert "Checking if solution vector satisfies formula: true" in lines
    _check_render(res, golden_dir / "aim100_sat.cnf")


def test_degenerate_replica_counts(golden_dir):
    """batch with zero replicas yields the reference's empty map (main.rs:276-308), inter refuses."""
    res = commands.batch(str(golden_dir / "aim100_sat.cnf"), 10, 0, step_size=0.01, log=lambda s: None)
    assert not res.is_satisfiable and res.values == {} and res.rendered == ""
    with pytest.raises(ValueError):
        commands.inter(str(golden_dir / "aim100_sat.cnf"), 0, step_number=10, step_size=0.01, log=lambda s: None)
    from odesat_b200 import batch as B
    from odesat_b200.system import DeviceFormula
    F = DeviceFormula(cnf.load_dimacs(str(golden_dir / "aim100_sat.cnf")))
    r = B.simulate_batch(F, 0, seed=1, step_size=0.01, steps=5, precision=L.F32)
    assert r.winner == -1 and r.steps_run == 0 and not r.assignment.any() and r.solved_step.size == 0
