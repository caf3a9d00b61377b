"""Implicit heat-equation time stepping with one multigrid solve per step (BASELINE config 5).

API-compatible with the reference's ``HeatSolver2D.solve_heat_problem(problem, nx, ny, time_config)``
(applications/heat_solver.py:117-262: same arguments, same result keys), but it solves the linear system
the reference's documentation states,

    (I - theta*alpha*dt*lap_h) u^{n+1} = u^n + (1-theta)*alpha*dt*lap_h u^n + dt*(theta f^{n+1} + (1-theta) f^n)

(docs/methodology.md:710; theta = 1 backward Euler, 1/2 Crank-Nicolson), divided by theta*alpha*dt:

    (-lap_h + lambda) u^{n+1} = lambda * rhs,     lambda = 1/(theta*alpha*dt),

with the Helmholtz-shifted multigrid cycle (`MixedPrecisionMultigrid(shift=lambda)`), instead of the reference's
fixed-point iteration around Poisson solves (heat_solver.py:308-390), which inherits the diverging default sign
(SURVEY fact 4).  There is therefore NO reference oracle for this row: parity is unpinned and the tests validate
against analytical solutions (O(dt) / O(dt^2) in time, O(h^2) in space) and the repo's own NumPy oracle.
The iterate stays in HBM between steps; the host sees it only at the end (or at save_frequency)."""
from __future__ import annotations

import time
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from .. import ops
from ..core.grid import Grid
from ..device import require_cuda, to_device, to_host
from ..problems import HeatProblem, TimeSteppingConfig, TimeSteppingMethod
from ..solvers.mixed_precision import MixedPrecisionMultigrid


class HeatSolver2D:
    def __init__(self, solver_type: str = "multigrid", max_levels: Optional[int] = None, max_iterations: int = 50,
                 tolerance: float = 1e-8, cycle_type: str = "V", use_gpu: bool = True, device_id: int = 0,
                 enable_mixed_precision: bool = True, precision_strategy: Optional[str] = None):
        if not use_gpu:
            raise ValueError("use_gpu=False is not available: this build has no CPU path")
        self.solver_type, self.max_levels, self.max_iterations = solver_type, max_levels, max_iterations
        self.tolerance, self.cycle_type, self.use_gpu, self.device_id = tolerance, cycle_type, use_gpu, device_id
        self.enable_mixed_precision = enable_mixed_precision
        self.precision_strategy = precision_strategy or ("adaptive" if enable_mixed_precision else "double")
        self.time_history: List[Dict[str, Any]] = []
        self._mg: Optional[MixedPrecisionMultigrid] = None
        self._mg_key = None

    @staticmethod
    def _theta(cfg: TimeSteppingConfig) -> float:
        if cfg.method == TimeSteppingMethod.BACKWARD_EULER:
            return 1.0
        if cfg.method == TimeSteppingMethod.CRANK_NICOLSON:
            return 0.5
        if cfg.method == TimeSteppingMethod.THETA_METHOD:
            if not 0.0 < cfg.theta <= 1.0:
                raise ValueError("theta must be in (0, 1] for an implicit step")
            return cfg.theta
        raise ValueError(f"Unknown time stepping method: {cfg.method}")

    def _solver(self, nx, ny, domain, lam, coefficient=None) -> MixedPrecisionMultigrid:
        key = (nx, ny, tuple(domain), lam, id(coefficient) if coefficient is not None else None)
        if self._mg is None or self._mg_key != key:
            self._mg = MixedPrecisionMultigrid(precision_strategy=self.precision_strategy, max_iterations=self.max_iterations,
                                               tolerance=self.tolerance, max_levels=self.max_levels,
                                               cycle_type=self.cycle_type, shift=lam, coefficient=coefficient,
                                               device=torch.device("cuda", self.device_id))
            self._mg.setup(nx, ny, domain)
            self._mg_key = key
        return self._mg

    def solve_heat_problem(self, problem: HeatProblem, nx: int, ny: int, time_config: TimeSteppingConfig,
                           save_solution_history: bool = False) -> Dict[str, Any]:
        """`problem.thermal_diffusivity`: a number alpha (u_t = alpha*lap u + f, the reference's problem class) or a
        callable a(X, Y) / an (nx, ny) array of nodal values (u_t = div(a grad u) + f, BASELINE configs[4]; README.md:175).
        Per step ONE kernel forms the right-hand side in the solver's own buffer (mg_heat_rhs: scaled sum, ring, norm)
        and the cycles run in place on the iterate, which never leaves the level-0 buffers of the cycle engine."""
        if self.precision_strategy in ("single", "fp32"):
            raise ValueError("the heat driver keeps the fp64 iterate between steps: use 'adaptive', 'refinement' or 'double'")
        dev = require_cuda(torch.device("cuda", self.device_id))
        domain = tuple(problem.domain)
        grid = Grid(nx, ny, domain)
        theta = self._theta(time_config)
        diff = problem.thermal_diffusivity
        variable = callable(diff) or isinstance(diff, (np.ndarray, torch.Tensor))
        if variable:  # -div(a grad u) + lambda u, lambda = 1/(theta dt); the diffusivity sits in the operator
            coef = np.array(np.broadcast_to(np.asarray(diff(grid.X, grid.Y) if callable(diff) else diff, dtype=np.float64),
                                            (nx, ny)))
            alpha = 1.0
        else:
            coef, alpha = None, float(diff)

        def source(t):
            if problem.source_function is None:
                return None
            f = np.asarray(problem.source_function(grid.X, grid.Y, t), dtype=np.float64)
            return None if not f.any() else to_device(np.array(np.broadcast_to(f, (nx, ny))), device=dev)[0]

        t_cur, dt, step = 0.0, time_config.dt, 0
        total_mg, solver_time = 0, 0.0
        mg = self._solver(nx, ny, domain, 1.0 / (theta * alpha * dt), coef)
        eng = mg._engine
        b = eng.levels[0].bufs(torch.float64)
        b.u.copy_(to_device(np.asarray(problem.initial_condition(grid.X, grid.Y), dtype=np.float64), device=dev)[0])
        ops.zero_ring_(b.u)
        time_steps, solutions = [0.0], [to_host(b.u)] if save_solution_history else []
        f_old = source(0.0) if theta < 1.0 else None
        t_start = time.time()
        while t_cur < time_config.t_final - 1e-14:
            if t_cur + dt > time_config.t_final:
                dt = time_config.t_final - t_cur
            step += 1
            t_new = t_cur + dt
            lam = 1.0 / (theta * alpha * dt)
            nmg = self._solver(nx, ny, domain, lam, coef)
            if nmg is not mg:  # shortened last step: another shift, hence another solver; hand the iterate over
                u_prev = eng.levels[0].bufs(torch.float64).u
                mg, eng = nmg, nmg._engine
                eng.levels[0].bufs(torch.float64).u.copy_(u_prev)
            t0 = time.time()
            b = eng.levels[0].bufs(torch.float64)
            a64 = mg._operator.coefficients(nx, ny, torch.float64) if variable else None
            f_new = source(t_new)
            # b.f = lambda * (u + (1-theta) dt L_h u + dt (theta f_new + (1-theta) f_old)), ring zeroed, norm: one kernel
            ss = ops.heat_rhs_(b.u, b.f, grid.hx, grid.hy, lam=lam, c_lap=(1.0 - theta) * alpha * dt, f1=f_new,
                               c_f1=dt * theta, f0=f_old if theta < 1.0 else None, c_f0=dt * (1.0 - theta), a=a64)
            # relative stopping test: the right-hand side scales with lambda
            scale = float(np.sqrt(grid.hx * grid.hy * ops.read_scalar(ss)))
            mg.tolerance = self.tolerance * max(scale, 1e-300)
            mg.switch_threshold = max(1e-6 * scale, mg.tolerance)
            t1 = time.time()
            _, info = mg._solve_device(True)  # cycles in place, from u^n
            solver_time += time.time() - t0
            self.rhs_time = getattr(self, "rhs_time", 0.0) + (t1 - t0)
            self.cycle_time = getattr(self, "cycle_time", 0.0) + info["cycle_time"]
            total_mg += info["iterations"]
            self.last_precisions = info["precision_history"]
            f_old, t_cur = f_new, t_new
            if save_solution_history and step % time_config.save_frequency == 0:
                time_steps.append(t_cur)
                solutions.append(to_host(eng.levels[0].bufs(torch.float64).u))
        u = eng.levels[0].bufs(torch.float64).u
        total = time.time() - t_start
        u_host = to_host(u)
        errors: Dict[str, Any] = {}
        results: Dict[str, Any] = {
            "problem_name": problem.name, "grid_size": (nx, ny),
            "time_config": {"method": time_config.method.value, "dt_initial": time_config.dt, "dt_final": dt,
                            "t_final": time_config.t_final, "adaptive_dt": time_config.adaptive_dt},
            "final_solution": u_host, "final_time": t_cur, "total_steps": step, "total_time": total,
            "total_solver_time": solver_time, "avg_mg_iterations": total_mg / step if step else 0,
            "total_mg_iterations": total_mg, "errors": errors, "solver_type": self.solver_type, "use_gpu": self.use_gpu,
        }
        if problem.analytical_solution is not None:
            exact = np.asarray(problem.analytical_solution(grid.X, grid.Y, t_cur), dtype=np.float64)
            err = u_host - exact
            l2e = float(np.sqrt(np.sum(err ** 2) * grid.hx * grid.hy))
            l2n = float(np.sqrt(np.sum(exact ** 2) * grid.hx * grid.hy))
            mx, mxn = float(np.max(np.abs(err))), float(np.max(np.abs(exact)))
            errors.update({"l2_error": l2e, "relative_l2_error": l2e / l2n if l2n > 0 else l2e, "max_error": mx,
                           "relative_max_error": mx / mxn if mxn > 0 else mx, "grid_spacing": (grid.hx, grid.hy)})
            results["analytical_solution"] = exact
        if save_solution_history:
            results["time_steps"] = np.array(time_steps)
            results["solution_history"] = solutions
        self.time_history.append(results)
        return results
