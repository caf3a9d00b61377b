"""BASELINE config 5: backward-Euler heat equation, pure diffusion, n x n grid, `steps` time steps, one shifted
multigrid solve per step.  Prints a JSON line (time per step, MG cycles per step, error vs analytical).

    python tools/heat_config5.py [n] [steps] [varcoef]                            one GPU (HeatSolver2D); `varcoef`:
                                                                                  u_t = div(a grad u), a = 1 + 0.5 sin(2 pi x) cos(pi y) + x y
                                                                                  (the variable-coefficient half of configs[4])
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        tools/heat_config5.py [n] [steps] [varcoef]                               N GPUs, row slabs (DistributedHeatSolver)
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mixed_precision_multigrid_solvers_for_pdes_b200 import (HeatSolver2D, HeatTestProblems, TimeSteppingConfig,  # noqa: E402
                                                             TimeSteppingMethod)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8193
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dt = 1e-4
prob = HeatTestProblems().get_problem("pure_diffusion")
prob.source_function = None  # identically zero: skip the per-step host evaluation
varcoef = len(sys.argv) > 3 and sys.argv[3] == "varcoef"
if varcoef:
    import numpy as np
    prob.thermal_diffusivity = lambda X, Y: 1.0 + 0.5 * np.sin(2 * np.pi * X) * np.cos(np.pi * Y) + X * Y
    prob.analytical_solution = None  # no closed form: the run is timed, accuracy is covered by tests/test_gpu_heat.py
cfg = TimeSteppingConfig(TimeSteppingMethod.BACKWARD_EULER, dt, dt * steps)
world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    import torch
    import torch.distributed as dist
    from mixed_precision_multigrid_solvers_for_pdes_b200.distributed import DistributedHeatSolver
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    s = DistributedHeatSolver(tolerance=1e-8, device=dev, use_cuda_graphs=True)
    s.solve_heat_problem(prob, n, n, TimeSteppingConfig(TimeSteppingMethod.BACKWARD_EULER, dt, dt * 8), gather=False)  # warm-up: graphs captured
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.time()
    res = s.solve_heat_problem(prob, n, n, cfg, gather=False)
    torch.cuda.synchronize()
    dist.barrier()
    wall = time.time() - t0
    rank = dist.get_rank()
else:
    s = HeatSolver2D(tolerance=1e-8)
    s.solve_heat_problem(prob, n, n, TimeSteppingConfig(TimeSteppingMethod.BACKWARD_EULER, dt, dt * 8))  # warm-up: graphs captured
    s.rhs_time = s.cycle_time = 0.0
    t0 = time.time()
    res = s.solve_heat_problem(prob, n, n, cfg)
    wall = time.time() - t0
    rank = 0
if rank == 0:
    print(json.dumps({"config": f"heat backward Euler {n}x{n}, {steps} steps, dt={dt}, {world} GPU(s)"
                                + (", variable diffusivity" if varcoef else ""), "n_gpus": world,
                      "wall_s": round(wall, 3), "solver_s": round(res["total_solver_time"], 3),
                      "ms_per_step": round(1e3 * res["total_solver_time"] / steps, 3),
                      "avg_mg_cycles_per_step": res["avg_mg_iterations"], "max_error": res["errors"].get("max_error"),
                      "relative_max_error": res["errors"].get("relative_max_error"),
                      "halo_exchanges": res.get("halo_exchanges"),
                      "rhs_s": round(getattr(s, "rhs_time", 0.0), 3), "cycles_s": round(getattr(s, "cycle_time", 0.0), 3),
                      "precision_history_last_step": getattr(s, "last_precisions", None)}))
if world > 1:
    del s, res
    import gc
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
