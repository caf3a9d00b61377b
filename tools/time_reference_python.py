#!/usr/bin/env python
"""Time the REFERENCE'S OWN Python classes on this path: MultigridSolver +
LaplacianOperator(-1) + GaussSeidelSmoother(red_black=True), V(2,2), tol 1e-8, fp64, coarsest grid 5 x 5, u0 = 0,
f = 2 pi^2 sin(pi x) sin(pi y) -- the oracle configuration of SURVEY 8c.

Runs in the build container only (imports /root/reference/src, which does not exist on the GPU box); the result is
committed as profiles/reference_python_timing.json and quoted by bench.py's `cpu_baseline.reference_python` next to
the C/OpenMP port, so the reader sees both the reference as shipped (single-threaded pure-Python loops) and the
fastest honest CPU restatement of it.

    python tools/time_reference_python.py [sizes...]      (default 129 257)
"""
import json
import logging
import os
import platform
import sys
import time

import numpy as np

sys.path.insert(0, "/root/reference/src")
logging.disable(logging.CRITICAL)

from multigrid.core.grid import Grid  # noqa: E402
from multigrid.operators.laplacian import LaplacianOperator  # noqa: E402
from multigrid.operators.transfer import ProlongationOperator, RestrictionOperator  # noqa: E402
from multigrid.solvers.multigrid import MultigridSolver  # noqa: E402
from multigrid.solvers.smoothers import GaussSeidelSmoother  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(n: int):
    L = int(np.log2(n - 1)) - 1
    g = Grid(n, n)
    op = LaplacianOperator(coefficient=-1.0)
    s = MultigridSolver(max_levels=L, max_iterations=50, tolerance=1e-8, cycle_type="V", pre_smooth_iterations=2,
                        post_smooth_iterations=2)
    s.setup(g, op, RestrictionOperator("full_weighting"), ProlongationOperator("bilinear"),
            smoother=GaussSeidelSmoother(red_black=True))
    f = 2 * np.pi ** 2 * np.sin(np.pi * g.X) * np.sin(np.pi * g.Y)
    t0 = time.perf_counter()
    u, info = s.solve(g, op, f)
    dt = time.perf_counter() - t0
    err = float(np.max(np.abs(u - np.sin(np.pi * g.X) * np.sin(np.pi * g.Y))))
    return {"n": n, "levels": L, "cycles": int(info["iterations"]), "final_residual": float(info["final_residual"]),
            "max_error": err, "seconds": dt, "unknowns_per_s": n * n * info["iterations"] / dt}


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [129, 257]
    rows = [run(n) for n in sizes]
    out = {"what": "the reference's own MultigridSolver (pure-Python loops, one thread), oracle configuration of SURVEY 8c",
           "where": f"build container, {platform.processor() or platform.machine()}, {os.cpu_count()} logical CPUs, "
                    f"Python {platform.python_version()}, NumPy {np.__version__}",
           "unit": "unknowns/s", "runs": rows,
           "value": float(np.mean([r["unknowns_per_s"] for r in rows]))}
    path = os.path.join(ROOT, "profiles", "reference_python_timing.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
