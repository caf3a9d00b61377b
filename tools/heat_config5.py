"""BASELINE config 5 on one GPU: backward-Euler heat equation, pure diffusion, n x n grid, `steps` time steps, one
shifted multigrid solve per step.  Prints a JSON line (time per step, MG cycles per step, error vs analytical)."""
import json
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mixed_precision_multigrid_solvers_for_pdes_b200 import (HeatSolver2D, HeatTestProblems, TimeSteppingConfig,  # noqa: E402
                                                             TimeSteppingMethod)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8193
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dt = 1e-4
prob = HeatTestProblems().get_problem("pure_diffusion")
prob.source_function = None  # identically zero: skip the per-step host evaluation
s = HeatSolver2D(tolerance=1e-8)
t0 = time.time()
res = s.solve_heat_problem(prob, n, n, TimeSteppingConfig(TimeSteppingMethod.BACKWARD_EULER, dt, dt * steps))
wall = time.time() - t0
print(json.dumps({"config": f"heat backward Euler {n}x{n}, {steps} steps, dt={dt}", "wall_s": round(wall, 3),
                  "solver_s": round(res["total_solver_time"], 3), "ms_per_step": round(1e3 * res["total_solver_time"] / steps, 3),
                  "avg_mg_cycles_per_step": res["avg_mg_iterations"], "max_error": res["errors"]["max_error"],
                  "relative_max_error": res["errors"]["relative_max_error"]}))
