#!/usr/bin/env python
"""Evidence for the rounding-floor stopping rule (solvers/policy.py): run the mixed-precision solve of BASELINE
configs[2] (16385^2, adaptive, switch at 1e-6) for a fixed number of cycles with every stopping rule off and record,
after each cycle, the h-scaled residual norm the driver sees and the MMS error max|u - sin(pi x) sin(pi y)| against the
closed-form discretisation error of the exactly converged discrete solution.

    python tools/floor_study.py [n] [cycles] > profiles/r02_floor_study_<n>.json
"""
import json
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mixed_precision_multigrid_solvers_for_pdes_b200 import MixedPrecisionMultigrid, ops  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16385
    cycles = int(sys.argv[2]) if len(sys.argv) > 2 else 13
    h = 1.0 / (n - 1)
    a = math.pi * h / 2.0
    want = (a / math.sin(a)) ** 2 - 1.0
    rows = []
    for strategy in ("adaptive", "refinement", "double"):
        s = MixedPrecisionMultigrid(strategy, switch_threshold=1e-6, tolerance=0.0, stop_on_rounding_floor=False,
                                    max_iterations=cycles)
        s.setup(n, n)
        b64 = s._engine.levels[0].bufs(torch.float64)
        ops.fill_sinsin_(b64.f, (0.0, 1.0, 0.0, 1.0), 2 * np.pi ** 2, 1.0, 1.0)
        ops.zero_ring_(b64.f)
        pol = s.make_policy()
        first = True
        if pol.phase == "refine":
            s._refinement_residual(u_zero=True)
        for k in range(1, cycles + 1):
            ph = pol.phase
            norm = s._cycle_refinement(u_zero=first, last_hint=pol.likely_last()) if ph == "refine" else s._cycle_fp64(u_zero=first)
            first = False
            pol.observe(norm)
            err = ops.maxerr_sinsin(s._engine.levels[0].bufs(torch.float64).u)
            rows.append({"strategy": strategy, "cycle": k, "phase": ph, "residual": norm, "max_error": err,
                         "error_vs_closed_form_percent": 100.0 * (err - want) / want})
        del s
        torch.cuda.empty_cache()
    print(json.dumps({"grid": n, "closed_form_error": want,
                      "fp64_floor_bound": 2.220446049250313e-16 * 4.0 / h ** 2 * 0.5, "cycles": rows}, indent=1))


if __name__ == "__main__":
    main()
