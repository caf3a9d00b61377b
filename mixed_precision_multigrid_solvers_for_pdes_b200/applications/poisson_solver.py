"""``PoissonSolver2D``: the application wrapper around the cycle, API-compatible with
applications/poisson_solver.py:35-313 of the reference (constructor arguments, ``solve_poisson_problem(problem, nx,
ny, initial_guess) -> dict`` with the same keys, ``_compute_errors`` norms).

Deliberate deviations, both documented reference defects:
  * the operator is ``LaplacianOperator(-1.0)``: the reference hard-codes ``LaplacianOperator()`` (+1)
    (poisson_solver.py:78), with which its own solver diverges by x2.3 per cycle;
  * the default smoother is red-black Gauss-Seidel (the kernel this build is about); ``smoother="lexicographic"``
    reproduces the reference's setup default (multigrid.py:112-117)."""
from __future__ import annotations

import time
from typing import Any, Dict, List, Optional

import numpy as np

from ..core.grid import Grid
from ..operators.laplacian import LaplacianOperator
from ..operators.transfer import ProlongationOperator, RestrictionOperator
from ..problems import PoissonProblem
from ..solvers.mixed_precision import MixedPrecisionMultigrid
from ..solvers.multigrid import MultigridSolver
from ..solvers.smoothers import GaussSeidelSmoother


def compute_errors(numerical: np.ndarray, analytical: np.ndarray, grid) -> Dict[str, Any]:
    """L2 / max / H1-seminorm errors exactly as poisson_solver.py:281-313."""
    error = numerical - analytical
    l2_error = np.sqrt(np.sum(error ** 2) * grid.hx * grid.hy)
    l2_norm = np.sqrt(np.sum(analytical ** 2) * grid.hx * grid.hy)
    max_error = np.max(np.abs(error))
    max_norm = np.max(np.abs(analytical))
    gx = np.diff(error, axis=0) / grid.hx
    gy = np.diff(error, axis=1) / grid.hy
    h1 = np.sqrt(np.sum(gx[:-1, :] ** 2) * grid.hx * grid.hy + np.sum(gy[:, :-1] ** 2) * grid.hx * grid.hy)
    return {"l2_error": float(l2_error), "relative_l2_error": float(l2_error / l2_norm if l2_norm > 0 else l2_error),
            "max_error": float(max_error), "relative_max_error": float(max_error / max_norm if max_norm > 0 else max_error),
            "h1_semi_error": float(h1), "grid_spacing": (grid.hx, grid.hy)}


class PoissonSolver2D:
    def __init__(self, solver_type: str = "multigrid", max_levels: int = 6, max_iterations: int = 100,
                 tolerance: float = 1e-8, cycle_type: str = "V", use_gpu: bool = True, device_id: int = 0,
                 enable_mixed_precision: bool = True, smoother: str = "red_black"):
        if not use_gpu:
            raise ValueError("use_gpu=False is not available: this build has no CPU path")
        self.solver_type, self.max_levels, self.max_iterations = solver_type, max_levels, max_iterations
        self.tolerance, self.cycle_type, self.use_gpu, self.device_id = tolerance, cycle_type, use_gpu, device_id
        self.enable_mixed_precision, self.smoother = enable_mixed_precision, smoother
        self.operator = LaplacianOperator(-1.0)
        self.restriction = RestrictionOperator("full_weighting")
        self.prolongation = ProlongationOperator("bilinear")
        self.solver = self._create_solver()
        self.current_problem: Optional[PoissonProblem] = None
        self.solve_history: List[Dict[str, Any]] = []

    def _create_solver(self):
        if self.solver_type in ("gpu_multigrid", "gpu_ca_multigrid", "mixed_precision") and self.enable_mixed_precision:
            return MixedPrecisionMultigrid("adaptive", max_iterations=self.max_iterations, tolerance=self.tolerance,
                                           max_levels=self.max_levels, cycle_type=self.cycle_type,
                                           device=f"cuda:{self.device_id}")
        return MultigridSolver(max_levels=self.max_levels, max_iterations=self.max_iterations, tolerance=self.tolerance,
                               cycle_type=self.cycle_type, device=f"cuda:{self.device_id}")

    def solve_poisson_problem(self, problem: PoissonProblem, nx: int, ny: int,
                              initial_guess: Optional[np.ndarray] = None) -> Dict[str, Any]:
        self.current_problem = problem
        grid = Grid(nx=nx, ny=ny, domain=tuple(problem.domain))
        rhs = np.asarray(problem.source_function(grid.X, grid.Y), dtype=np.float64)
        bc = problem.boundary_conditions or {"type": "dirichlet", "value": 0.0}
        if bc.get("type", "dirichlet") != "dirichlet" or bc.get("value", 0.0) not in (0, 0.0):
            raise NotImplementedError("the hot path implements homogeneous Dirichlet boundary data only "
                                      "(so does the reference: poisson_solver.py:203-207 cannot apply anything else)")
        t0 = time.time()
        if isinstance(self.solver, MixedPrecisionMultigrid):
            solution, info = self.solver.solve(PoissonProblem(problem.name, rhs=rhs, nx=nx, ny=ny, domain=problem.domain),
                                               initial_guess=initial_guess)
            solution = np.array(solution)
        else:
            sm = GaussSeidelSmoother(red_black=True) if self.smoother == "red_black" else None
            self.solver.setup(grid, self.operator, self.restriction, self.prolongation, smoother=sm)
            solution, info = self.solver.solve(grid, self.operator, rhs, initial_guess)
        solve_time = time.time() - t0
        results: Dict[str, Any] = {
            "problem_name": problem.name, "grid_size": (nx, ny), "domain": problem.domain, "solution": solution,
            "solve_time": solve_time, "solver_info": info, "errors": {}, "solver_type": self.solver_type,
            "use_gpu": self.use_gpu, "mixed_precision": self.enable_mixed_precision,
        }
        if problem.analytical_solution is not None:
            exact = np.asarray(problem.analytical_solution(grid.X, grid.Y), dtype=np.float64)
            results["errors"] = compute_errors(solution, exact, grid)
            results["analytical_solution"] = exact
        self.solve_history.append(results)
        return results

    def run_convergence_study(self, problem: PoissonProblem, grid_sizes=(17, 33, 65, 129)) -> Dict[str, Any]:
        """Errors on a sequence of grids and the observed order (poisson_solver.py:315-396, condensed)."""
        rows = [self.solve_poisson_problem(problem, n, n) for n in grid_sizes]
        hs = np.array([r["errors"]["grid_spacing"][0] for r in rows])
        out: Dict[str, Any] = {"grid_sizes": list(grid_sizes), "results": rows}
        for key in ("l2_error", "max_error"):
            e = np.array([r["errors"][key] for r in rows])
            out[f"{key}_rate"] = float(np.polyfit(np.log(hs), np.log(e), 1)[0])
        return out

    def benchmark_solver_performance(self, problem: PoissonProblem, grid_sizes, num_runs: int = 5) -> Dict[str, Any]:
        """Wall-clock statistics of repeated solves per grid size (poisson_solver.py:398-460: same arguments, same
        result keys; `throughput` = unknowns per second of a whole solve, as there)."""
        rows = []
        for nx, ny in grid_sizes:
            times, iters = [], []
            for _ in range(num_runs):
                r = self.solve_poisson_problem(problem, nx, ny)
                times.append(r["solve_time"])
                iters.append(r["solver_info"]["iterations"])
            avg = float(np.mean(times))
            rows.append({"grid_size": (nx, ny), "total_unknowns": nx * ny, "num_runs": num_runs, "avg_time": avg,
                         "std_time": float(np.std(times)), "min_time": float(min(times)), "max_time": float(max(times)),
                         "avg_iterations": float(np.mean(iters)), "throughput": nx * ny / avg,
                         "solver_type": self.solver_type})
        return {"problem_name": problem.name,
                "solver_configuration": {"solver_type": self.solver_type, "use_gpu": self.use_gpu,
                                         "mixed_precision": self.enable_mixed_precision, "max_levels": self.max_levels,
                                         "cycle_type": self.cycle_type},
                "benchmark_results": rows}

    def get_solver_statistics(self) -> Dict[str, Any]:
        """poisson_solver.py:462-480."""
        stats: Dict[str, Any] = {
            "solver_type": self.solver_type,
            "configuration": {"max_levels": self.max_levels, "max_iterations": self.max_iterations,
                              "tolerance": self.tolerance, "cycle_type": self.cycle_type, "use_gpu": self.use_gpu,
                              "mixed_precision": self.enable_mixed_precision},
            "solve_history_count": len(self.solve_history)}
        if hasattr(self.solver, "get_performance_statistics"):
            stats["performance_statistics"] = self.solver.get_performance_statistics()
        return stats
