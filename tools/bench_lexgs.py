"""Times the lexicographic Gauss-Seidel sweep (mg_smooth_lexgs, the reference's default smoother) per grid size and a
MultigridSolver solve with the reference's defaults (lexicographic GS on every level) beside the red-black one.

    python tools/bench_lexgs.py
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mixed_precision_multigrid_solvers_for_pdes_b200 import (GaussSeidelSmoother, Grid, LaplacianOperator,  # noqa: E402
                                                             MultigridSolver, ProlongationOperator,
                                                             RestrictionOperator, ops)
from mixed_precision_multigrid_solvers_for_pdes_b200.device import empty_field  # noqa: E402

for n in (129, 1025, 4097, 16385):
    for dt in (torch.float64, torch.float32):
        u, f = empty_field(n, n, dt), empty_field(n, n, dt)
        u.zero_()
        f.fill_(1.0)
        h = 1.0 / (n - 1)
        ops.smooth_lexgs_(u, f, h, h, 1.0, 1)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        a.record()
        ops.smooth_lexgs_(u, f, h, h, 1.0, reps)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        print(json.dumps({"kernel": "lexgs sweep", "n": n, "dtype": str(dt).split(".")[1], "ms_per_sweep": round(ms, 4),
                          "points_per_us": round(n * n / ms / 1e3, 1)}), flush=True)

for n in (1025, 4097):
    for name, sm in (("lexicographic (reference default)", None), ("red-black", GaussSeidelSmoother(red_black=True))):
        g = Grid(n, n)
        x = np.linspace(0, 1, n)
        X, Y = np.meshgrid(x, x, indexing="ij")
        rhs = 2 * np.pi ** 2 * np.sin(np.pi * X) * np.sin(np.pi * Y)
        s = MultigridSolver(max_levels=20, tolerance=1e-8)
        s.setup(g, LaplacianOperator(-1.0), RestrictionOperator(), ProlongationOperator(), smoother=sm)
        s.solve(g, LaplacianOperator(-1.0), rhs)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, info = s.solve(g, LaplacianOperator(-1.0), rhs)
        torch.cuda.synchronize()
        print(json.dumps({"solve": f"MultigridSolver V(2,2) {n}x{n} fp64, smoother: {name}", "iterations": info["iterations"],
                          "seconds": round(time.perf_counter() - t0, 4), "final_residual": info["final_residual"]}), flush=True)
