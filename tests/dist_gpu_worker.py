"""torchrun worker of tests/test_gpu_distributed.py: solve a 1025 x 513 problem on WORLD_SIZE GPUs, rank 0 saves."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mixed_precision_multigrid_solvers_for_pdes_b200.distributed import DistributedMixedPrecisionSolver  # noqa: E402
from oracle import np_oracle as O  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    nx, ny, dom = 1025, 513, (0.0, 2.0, 0.0, 1.0)
    g = O.OGrid(nx, ny, dom)
    x, y = g.coords()
    f = 2 * np.pi ** 2 * np.sin(np.pi * x)[:, None] * np.sin(np.pi * y)[None, :]
    import time
    res = {}
    t00 = time.time()
    p2p = os.environ.get("MGB200_TEST_HALO") == "p2p"  # ghost rows pushed over peer memory (halo.py) instead of NCCL
    for strategy in ("double", "adaptive", "adaptive_graphs"):
        kw = {}
        if p2p:
            from mixed_precision_multigrid_solvers_for_pdes_b200.distributed import GHOST
            from mixed_precision_multigrid_solvers_for_pdes_b200.halo import SymmMemTransport
            kw["transport"] = SymmMemTransport(dev, GHOST)
        sol = DistributedMixedPrecisionSolver(nx, ny, domain=dom, precision_strategy=strategy.split("_")[0],
                                              use_fused_defect_down="always",
                                              tolerance=1e-8, agglomerate_below=129, device=dev,
                                              use_cuda_graphs=strategy.endswith("graphs"), **kw)
        sol.set_rhs_from_global(torch.from_numpy(f).to(dev))
        print(f"[worker r{dist.get_rank()}] {strategy}: start solve at {time.time() - t00:.1f}s D={sol.eng.D}", file=sys.stderr, flush=True)
        u, info = sol.solve()
        print(f"[worker r{dist.get_rank()}] {strategy}: solve done at {time.time() - t00:.1f}s it={info['iterations']} ex={sol.eng.exchanges}", file=sys.stderr, flush=True)
        if strategy.endswith("graphs"):  # second and third solve replay captured graphs (NCCL send/recv included)
            u, info = sol.solve()
            u, info = sol.solve()
            res["graphs_captured"] = sol.graphs.captured
        if dist.get_rank() == 0:
            print(f"[worker] {strategy}: solved in {time.time() - t00:.1f}s since start, {info['iterations']} cycles, "
                  f"graphs={sol.graphs.captured}", file=sys.stderr, flush=True)
        full = sol.eng.gather_solution(u)
        res[strategy] = {"iterations": info["iterations"], "history": info["residual_history"],
                         "u": full.cpu().numpy()}
        res["exchanges"] = info["halo_exchanges"]
        res["D"] = sol.eng.D
    if p2p:
        res["pushes"] = kw["transport"].pushes
        if dist.get_rank() == 0:
            torch.save(res, sys.argv[1])
        del sol, u, full, kw
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        return
    # BASELINE configs[4] on slabs: 3 backward-Euler steps + a shortened 4th, one shifted solve per step, graphs on
    from mixed_precision_multigrid_solvers_for_pdes_b200 import HeatProblem, TimeSteppingConfig, TimeSteppingMethod
    from mixed_precision_multigrid_solvers_for_pdes_b200.distributed import DistributedHeatSolver
    kx, ky = np.pi / 2.0, np.pi
    mode = lambda X, Y: np.sin(kx * X) * np.sin(ky * Y)  # noqa: E731
    prob = HeatProblem("decay", mode, None, lambda X, Y, t: mode(X, Y) * np.exp(-(kx ** 2 + ky ** 2) * t), domain=dom)
    hs = DistributedHeatSolver(tolerance=1e-9, agglomerate_below=129, device=dev, use_cuda_graphs=True,
                               use_fused_defect_down="always")
    r = hs.solve_heat_problem(prob, nx, ny, TimeSteppingConfig(TimeSteppingMethod.BACKWARD_EULER, dt=0.002, t_final=0.007))
    res["heat"] = {"u": r["final_solution"], "iters": r["total_mg_iterations"], "steps": r["total_steps"],
                   "errors": r["errors"], "exchanges": r["halo_exchanges"]}
    del hs, r
    if dist.get_rank() == 0:
        torch.save(res, sys.argv[1])
    # captured graphs hold NCCL work: release them before tearing the communicator down
    del sol, u, full
    import gc
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
