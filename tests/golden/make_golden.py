"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports the unmodified reference package from /root/reference/src (multigrid.core,
multigrid.operators, multigrid.solvers need only NumPy) and records, for seeded random
inputs and for the manufactured sin(pi x) sin(pi y) problem, the outputs of

  * LaplacianOperator.apply / .residual          (operators/laplacian.py:44,105)
  * GaussSeidelSmoother (red-black / lexicographic), JacobiSmoother, WeightedJacobiSmoother
                                                  (solvers/smoothers.py:41,117,210)
  * RestrictionOperator / ProlongationOperator    (operators/transfer.py:53,189)
  * Grid.l2_norm                                  (core/grid.py:174)
  * MultigridSolver.solve full runs (V/W/F, several smoothers, PrecisionManager('mixed'))
                                                  (solvers/multigrid.py:184)

into ``ops_golden.npz`` and ``solve_golden.npz``.  The reference has no golden vectors of its own
(SURVEY.md section 4), so these runs are what pins the oracle and the CUDA path.
NumPy version matters only for the float32 cases (NEP 50 scalar promotion); it is recorded.
"""
import json
import logging
import os
import sys
import time

import numpy as np

REF = "/root/reference/src"
sys.path.insert(0, REF)
logging.disable(logging.CRITICAL)

from multigrid.core.grid import Grid  # noqa: E402
from multigrid.core.precision import PrecisionManager  # noqa: E402
from multigrid.operators.laplacian import LaplacianOperator  # noqa: E402
from multigrid.operators.transfer import ProlongationOperator, RestrictionOperator  # noqa: E402
from multigrid.solvers.multigrid import MultigridSolver  # noqa: E402
from multigrid.solvers.smoothers import (GaussSeidelSmoother, JacobiSmoother,  # noqa: E402
                                         WeightedJacobiSmoother)

HERE = os.path.dirname(os.path.abspath(__file__))

# (nx, ny, domain): square power-of-two, non-square (hx != hy), non-unit domain, smallest legal
OP_CASES = [
    (9, 9, (0.0, 1.0, 0.0, 1.0)),
    (17, 33, (0.0, 1.0, 0.0, 1.0)),
    (33, 17, (0.0, 2.0, 0.0, 1.0)),
    (5, 5, (0.0, 1.0, 0.0, 1.0)),
    (13, 21, (-1.0, 2.0, 0.5, 1.7)),   # h not a power of two
    (65, 65, (0.0, 1.0, 0.0, 1.0)),
]


def ops_fixtures():
    out = {}
    meta = []
    rng = np.random.default_rng(20261018)
    for ci, (nx, ny, dom) in enumerate(OP_CASES):
        for dt in (np.float64, np.float32):
            g = Grid(nx, ny, dom, dtype=dt)
            u = rng.uniform(-1, 1, (nx, ny)).astype(dt)
            f = rng.uniform(-1, 1, (nx, ny)).astype(dt)
            key = f"c{ci}_{np.dtype(dt).name}"
            out[f"{key}_u"] = u
            out[f"{key}_f"] = f
            for coeff in (1.0, -1.0, 2.5):
                op = LaplacianOperator(coeff)
                out[f"{key}_apply_{coeff}"] = op.apply(g, u)
                out[f"{key}_residual_{coeff}"] = op.residual(g, u, f)
            out[f"{key}_l2"] = np.array(g.l2_norm(f))
            for omega in (1.0, 1.3):
                for sweeps in (1, 2, 3):
                    out[f"{key}_rbgs_{omega}_{sweeps}"] = GaussSeidelSmoother(
                        relaxation_parameter=omega, red_black=True).smooth(g, None, u, f, sweeps)
                out[f"{key}_lexgs_{omega}_2"] = GaussSeidelSmoother(
                    relaxation_parameter=omega, red_black=False).smooth(g, None, u, f, 2)
            for sweeps in (1, 3):
                out[f"{key}_jacobi_{sweeps}"] = JacobiSmoother().smooth(g, None, u, f, sweeps)
                out[f"{key}_wjacobi_{sweeps}"] = WeightedJacobiSmoother().smooth(g, None, u, f, sweeps)
            cg = g.coarsen()
            for m in ("full_weighting", "injection", "half_weighting"):
                out[f"{key}_restrict_{m}"] = RestrictionOperator(m).apply(g, u, cg)
            uc = rng.uniform(-1, 1, cg.shape).astype(dt)
            out[f"{key}_uc"] = uc
            for m in ("bilinear", "injection"):
                out[f"{key}_prolong_{m}"] = ProlongationOperator(m).apply(cg, uc, g)
            meta.append({"key": key, "nx": nx, "ny": ny, "domain": dom, "dtype": np.dtype(dt).name})
    # dtype-mixing probe (SURVEY 8a behaviour probes): fp32 field, fp64 coarse grid -> fp64 result
    g64 = Grid(17, 17)
    u32 = rng.uniform(-1, 1, (17, 17)).astype(np.float32)
    out["probe_restrict_f32_in_f64_grid_u"] = u32
    out["probe_restrict_f32_in_f64_grid"] = RestrictionOperator().apply(g64, u32, g64.coarsen())
    return out, meta


def mms_rhs(g):
    return 2 * np.pi ** 2 * np.sin(np.pi * g.X) * np.sin(np.pi * g.Y)


SOLVE_CASES = [
    # name, nx, ny, max_levels, cycle, smoother, dtype, precision
    ("v33", 33, 33, 4, "V", "rbgs", "float64", None),
    ("v65_l4", 65, 65, 4, "V", "rbgs", "float64", None),
    ("v65", 65, 65, 5, "V", "rbgs", "float64", None),
    ("v129", 129, 129, 6, "V", "rbgs", "float64", None),
    ("v129_l4", 129, 129, 4, "V", "rbgs", "float64", None),
    ("w129", 129, 129, 6, "W", "rbgs", "float64", None),
    ("w65", 65, 65, 5, "W", "rbgs", "float64", None),
    ("f65", 65, 65, 5, "F", "rbgs", "float64", None),
    ("v65_jacobi", 65, 65, 5, "V", "jacobi", "float64", None),
    ("v65_wjacobi", 65, 65, 5, "V", "wjacobi", "float64", None),
    ("v65_lexgs", 65, 65, 5, "V", "lexgs", "float64", None),
    ("v33x65", 33, 65, 4, "V", "rbgs", "float64", None),
    ("v65_mixed_levels", 65, 65, 5, "V", "rbgs", "float64", "mixed"),
    ("v65_f32", 65, 65, 5, "V", "rbgs", "float32", None),
    ("v257", 257, 257, 7, "V", "rbgs", "float64", None),
]


def make_smoother(kind):
    if kind == "rbgs":
        return GaussSeidelSmoother(red_black=True)
    if kind == "lexgs":
        return None  # MultigridSolver.setup default (multigrid.py:112-117)
    if kind == "jacobi":
        return JacobiSmoother()
    if kind == "wjacobi":
        return WeightedJacobiSmoother()
    raise ValueError(kind)


def solve_fixtures():
    out = {}
    meta = []
    for name, nx, ny, L, cyc, sm, dt, prec in SOLVE_CASES:
        t0 = time.time()
        g = Grid(nx, ny, dtype=np.dtype(dt).type)
        op = LaplacianOperator(-1.0)
        s = MultigridSolver(max_levels=L, max_iterations=30 if dt == "float64" else 12, tolerance=1e-8,
                            cycle_type=cyc, pre_smooth_iterations=2, post_smooth_iterations=2)
        s.setup(g, op, RestrictionOperator("full_weighting"), ProlongationOperator("bilinear"),
                smoother=make_smoother(sm))
        f = mms_rhs(g).astype(np.dtype(dt).type)
        pm = PrecisionManager(default_precision=prec, adaptive=True) if prec else None
        u, info = s.solve(g, op, f, precision_manager=pm)
        exact = np.sin(np.pi * g.X) * np.sin(np.pi * g.Y)
        out[f"{name}_hist"] = np.array(info["residual_history"], dtype=np.float64)
        if nx * ny <= 129 * 129:
            out[f"{name}_u"] = u
        meta.append({"name": name, "nx": nx, "ny": ny, "max_levels": L, "cycle": cyc, "smoother": sm,
                     "dtype": dt, "precision": prec, "iterations": int(info["iterations"]),
                     "converged": bool(info["converged"]), "num_levels": int(info["num_levels"]),
                     "max_error": float(np.max(np.abs(u.astype(np.float64) - exact))),
                     "u_dtype": str(u.dtype), "seconds": round(time.time() - t0, 2)})
        print(meta[-1])
    return out, meta


def main():
    ops, ops_meta = ops_fixtures()
    np.savez_compressed(os.path.join(HERE, "ops_golden.npz"), **ops)
    sol, sol_meta = solve_fixtures()
    np.savez_compressed(os.path.join(HERE, "solve_golden.npz"), **sol)
    with open(os.path.join(HERE, "golden_meta.json"), "w") as fh:
        json.dump({"numpy": np.__version__, "reference": REF, "ops": ops_meta, "solves": sol_meta}, fh, indent=1)
    print("wrote", len(ops), "op arrays,", len(sol), "solve arrays")


if __name__ == "__main__":
    main()
