"""Golden runs of the reference's CorrectedMultigridSolver (solvers/corrected_multigrid.py), the solver its validation
modules and tutorials run.  Build container only (needs /root/reference):

    python tests/golden/make_golden_corrected.py      ->  tests/golden/corrected_golden.npz"""
import json
import logging
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference/src")
logging.disable(logging.CRITICAL)

from multigrid.core.grid import Grid  # noqa: E402
from multigrid.solvers.corrected_multigrid import CorrectedMultigridSolver  # noqa: E402

CASES = [
    dict(name="mms33_L4", n=33, levels=4, problem="manufactured", max_iterations=20, tolerance=1e-10),
    dict(name="mms65_L4", n=65, levels=4, problem="manufactured", max_iterations=6, tolerance=1e-10),
    dict(name="poly17_L3", n=17, levels=3, problem="polynomial", max_iterations=15, tolerance=1e-9),
    dict(name="mms33_L6", n=33, levels=6, problem="manufactured", max_iterations=8, tolerance=1e-8),
]


def main():
    out, meta = {}, []
    for c in CASES:
        g = Grid(c["n"], c["n"], domain=(0, 1, 0, 1))
        s = CorrectedMultigridSolver(max_levels=c["levels"], max_iterations=c["max_iterations"], tolerance=c["tolerance"])
        rhs, exact = s.create_test_problem(g, c["problem"])
        r = s.solve(np.zeros_like(rhs), rhs, g)
        out[c["name"] + "_rhs"] = rhs
        out[c["name"] + "_u"] = r["solution"]
        out[c["name"] + "_hist"] = np.array(r["residual_history"])
        meta.append(dict(c, iterations=int(r["iterations"]), converged=bool(r["converged"]),
                         hierarchy=[list(gr.shape) for gr in s.grids],
                         max_error=float(np.max(np.abs((r["solution"] - exact)[1:-1, 1:-1])))))
        print(c["name"], r["iterations"], r["converged"], "%.3e" % r["final_residual"], meta[-1]["hierarchy"], "%.6e" % meta[-1]["max_error"])
    out["meta"] = np.array(json.dumps({"numpy": np.__version__, "cases": meta}))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "corrected_golden.npz"), **out)


if __name__ == "__main__":
    main()
