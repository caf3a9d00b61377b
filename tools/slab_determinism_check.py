"""Decomposition-independence check on real GPUs: the slab solvers with world = N (torchrun) against the same
solvers on a private 1-rank group, in ONE job -- heat stepping (sin and polynomial initial data, eager and CUDA
graphs) and shifted / unshifted Poisson solves from random data.  Prints the max |difference| (must be 0).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/slab_determinism_check.py

History: with a Helmholtz shift the two differed in the last bit until every rounding of the point update was
pinned (nvcc fused `* 1/diag` into a later add in the unmasked fast path only); see mg_stream.cuh relax_fast."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mixed_precision_multigrid_solvers_for_pdes_b200 import HeatProblem, TimeSteppingConfig, TimeSteppingMethod
from mixed_precision_multigrid_solvers_for_pdes_b200 import distributed as D

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank = dist.get_rank()
solo = None
for r in range(dist.get_world_size()):
    g = dist.new_group([r])
    if r == rank:
        solo = g
nx, ny, dom = 1025, 513, (0.0, 2.0, 0.0, 1.0)
kx, ky = np.pi / 2.0, np.pi
mode = lambda X, Y: np.sin(kx * X) * np.sin(ky * Y)
prob = HeatProblem("decay", mode, None, lambda X, Y, t: mode(X, Y) * np.exp(-(kx ** 2 + ky ** 2) * t), domain=dom)
infos = []
orig = D.DistributedMixedPrecisionSolver.solve
def rec(self, keep_iterate=False):
    u, info = orig(self, keep_iterate)
    infos.append((len(info["residual_history"]), [s["iteration"] for s in info["precision_switches"]], info["residual_history"][-1], self.tolerance))
    return u, info
D.DistributedMixedPrecisionSolver.solve = rec
poly = HeatProblem("poly", lambda X, Y: X * (2.0 - X) * Y * (1.0 - Y) * (1.0 + X), None, None, domain=dom)
for strategy, graphs, tf, pr in (("double", False, 0.002, prob), ("double", False, 0.002, poly), ("adaptive", True, 0.007, prob), ("adaptive", True, 0.007, poly)):
    out = {}
    for name, grp in (("w2", None), ("w1", solo)):
        infos.clear()
        hs = D.DistributedHeatSolver(tolerance=1e-9, agglomerate_below=129, device=dev, use_cuda_graphs=graphs,
                                     precision_strategy=strategy, group=grp)
        r = hs.solve_heat_problem(pr, nx, ny, TimeSteppingConfig(TimeSteppingMethod.BACKWARD_EULER, dt=0.002, t_final=tf))
        out[name] = (r["final_solution"], list(infos), r["errors"].get("max_error", 0.0))
        del hs
    if rank == 0:
        d = np.abs(out["w2"][0] - out["w1"][0])
        rows = np.nonzero(d.max(axis=1))[0]
        print(f"== {pr.name} {strategy} graphs={graphs} t_final={tf}: maxdiff={d.max():.3e} nrows_diff={len(rows)} "
              f"rows[{rows.min() if len(rows) else -1}..{rows.max() if len(rows) else -1}] err w2={out['w2'][2]:.6e} w1={out['w1'][2]:.6e}", flush=True)
        for a, b in zip(out["w2"][1], out["w1"][1]):
            print("   w2", a, "| w1", b, flush=True)
rng = np.random.default_rng(3)
fglob = rng.uniform(-1, 1, (nx, ny)); fglob[0] = fglob[-1] = 0; fglob[:, 0] = fglob[:, -1] = 0
u0 = rng.uniform(-1, 1, (nx, ny)); u0[0] = u0[-1] = 0; u0[:, 0] = u0[:, -1] = 0
for strategy, sh in (("double", 0.0), ("double", 500.0), ("adaptive", 0.0)):
    out = {}
    for name, grp in (("w2", None), ("w1", solo)):
        sol = D.DistributedMixedPrecisionSolver(nx, ny, domain=dom, precision_strategy=strategy, tolerance=1e-7, max_iterations=4,
                                                agglomerate_below=129, device=dev, group=grp, shift=sh)
        sol.set_rhs_from_global(torch.from_numpy(fglob).to(dev))
        s0 = sol.s0
        b = sol.eng.bufs(0, torch.float64)
        b.u.copy_(torch.from_numpy(u0[s0.row0:s0.row0 + s0.loc_nx]).to(dev))
        sol.eng.set_valid(b.u, sol.eng.part.ghost)
        u, info = orig(sol, True)
        out[name] = (sol.eng.gather_solution(u).cpu().numpy(), info["residual_history"])
    if rank == 0:
        d = np.abs(out["w2"][0] - out["w1"][0])
        print(f"== poisson-random {strategy} shift={sh}: maxdiff={d.max():.3e}  hist w2={out['w2'][1]} w1={out['w1'][1]}", flush=True)
torch.cuda.synchronize()
dist.barrier()
dist.destroy_process_group()
