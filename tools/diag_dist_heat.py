import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from mixed_precision_multigrid_solvers_for_pdes_b200 import HeatTestProblems, TimeSteppingConfig, TimeSteppingMethod
from mixed_precision_multigrid_solvers_for_pdes_b200.distributed import DistributedHeatSolver
prob = HeatTestProblems().get_problem("pure_diffusion"); prob.source_function = None
n = 2049; dt = 1e-4
s = DistributedHeatSolver(tolerance=1e-8, device=torch.device("cuda", 0), use_cuda_graphs=True)
for steps in (3, 10):
    t0 = time.time()
    r = s.solve_heat_problem(prob, n, n, TimeSteppingConfig(TimeSteppingMethod.BACKWARD_EULER, dt, dt * steps), gather=False)
    torch.cuda.synchronize()
    sol = list(s._solvers.values())[0]
    print("steps", steps, "wall", round(time.time() - t0, 3), "solver_s", round(r["total_solver_time"], 4), "ms/step", round(1e3 * r["total_solver_time"] / steps, 3),
          "graphs", sol.graphs.captured, "entries", len(sol.graphs.entries), flush=True)
