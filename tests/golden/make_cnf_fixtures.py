"""Regenerates the DIMACS fixtures under tests/golden/ from the reference's three instance
files (public DIMACS `aim-100-1_6` instances + a 3-clause toy).  Runs only in the build container
(reads /root/reference/tests); the emitted files are committed because /root/reference does
not exist on the GPU box.  Comment headers are dropped, clause lines are kept verbatim."""
import sys
from pathlib import Path

SRC = Path("/root/reference/tests")
DST = Path(__file__).resolve().parent
NAMES = {"easy.cnf": "aim100_sat.cnf", "hard.cnf": "aim100_unsat.cnf", "small.cnf": "toy_mixed.cnf"}

for src, dst in NAMES.items():
    text = (SRC / src).read_text()
    keep_trailing_nl = text.endswith("\n")
    lines = [l for l in text.split("\n") if not l.startswith("c")]
    body = "\n".join(l for l in lines if l != "" or False)
    (DST / dst).write_text(body + ("\n" if keep_trailing_nl else ""))
    print(dst, len(body.splitlines()), "lines")
