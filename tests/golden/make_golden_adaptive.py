"""Golden runs of the reference's AdaptivePrecisionSolver (solvers/iterative.py:379-552), the second statement of the
precision-switching rule.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_adaptive.py      ->  tests/golden/adaptive_golden.json

Cases: the wrapper around JacobiSmoother / GaussSeidelSmoother(red_black=True) on 17x17 and 33x33 manufactured problems
(coefficient -1), with a PrecisionManager(adaptive=True) and without one, default and tightened switch rules."""
import json
import logging
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference/src")
logging.disable(logging.CRITICAL)

from multigrid.core.grid import Grid  # noqa: E402
from multigrid.core.precision import PrecisionManager  # noqa: E402
from multigrid.operators.laplacian import LaplacianOperator  # noqa: E402
from multigrid.solvers.iterative import AdaptivePrecisionSolver  # noqa: E402
from multigrid.solvers.smoothers import GaussSeidelSmoother, JacobiSmoother  # noqa: E402

CASES = [
    dict(name="jacobi17_pm", n=17, base="jacobi", max_iterations=40, tolerance=1e-3, pm=True, threshold=0.95, window=5, min_it=10),
    dict(name="rbgs17_pm", n=17, base="rbgs", max_iterations=60, tolerance=3.0, pm=True, threshold=0.9, window=3, min_it=5),
    dict(name="jacobi33_pm_never", n=33, base="jacobi", max_iterations=25, tolerance=1e-9, pm=True, threshold=1.5, window=5, min_it=10),
    dict(name="rbgs17_nopm", n=17, base="rbgs", max_iterations=30, tolerance=1e-4, pm=False, threshold=0.95, window=5, min_it=10),
]


def main():
    out = {"numpy": np.__version__, "cases": []}
    for c in CASES:
        g = Grid(c["n"], c["n"])
        op = LaplacianOperator(coefficient=-1.0)
        f = 2 * np.pi ** 2 * np.sin(np.pi * g.X) * np.sin(np.pi * g.Y)
        base = (JacobiSmoother(max_iterations=c["max_iterations"], tolerance=c["tolerance"]) if c["base"] == "jacobi" else
                GaussSeidelSmoother(max_iterations=c["max_iterations"], tolerance=c["tolerance"], red_black=True))
        s = AdaptivePrecisionSolver(base, precision_switch_threshold=c["threshold"], convergence_window=c["window"],
                                    min_iterations_before_switch=c["min_it"])
        pm = PrecisionManager(default_precision="double", adaptive=True) if c["pm"] else None
        u, info = s.solve(g, op, f, precision_manager=pm)
        out["cases"].append(dict(c, iterations=info["iterations"], converged=bool(info["converged"]),
                                 residual_history=[float(x) for x in info["residual_history"]],
                                 precision_switched=bool(info["precision_switched"]),
                                 switch_iteration=info["switch_iteration"],
                                 precision_levels=list(s.history.precision_levels),
                                 final_precision=(pm.current_precision.value if pm else None),
                                 u_dtype=str(u.dtype), u_sum=float(np.sum(u)), u_max=float(np.max(np.abs(u))),
                                 solver_name=s.name, keys=sorted(info.keys())))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "adaptive_golden.json")
    json.dump(out, open(path, "w"), indent=1)
    for c in out["cases"]:
        print(c["name"], c["iterations"], c["converged"], c["precision_switched"], c["switch_iteration"], c["final_precision"],
              c["u_dtype"], set(c["precision_levels"]))


if __name__ == "__main__":
    main()
