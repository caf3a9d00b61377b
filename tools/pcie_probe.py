"""What bounds the end-to-end (host buffers) solve: raw PCIe rates of this box beside the solve_many() pipeline.
Prints JSON lines: H2D alone, D2H alone, both at once, each and both under a running solve, and solve_many batches.

    python tools/pcie_probe.py [n]
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mixed_precision_multigrid_solvers_for_pdes_b200 import MixedPrecisionMultigrid, PoissonProblem, ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16385
dev = torch.device("cuda", 0)
h_in = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
h_out = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
h_in.fill_(1.0)
d_in = torch.empty((n, n), dtype=torch.float64, device=dev)
d_out = torch.zeros((n, n), dtype=torch.float64, device=dev)
gb = n * n * 8 / 1e9
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d()
    d2h()


t = timed(h2d)
print(json.dumps({"what": "H2D alone", "ms": round(t * 1e3, 2), "GB/s": round(gb / t, 1)}), flush=True)
t = timed(d2h)
print(json.dumps({"what": "D2H alone", "ms": round(t * 1e3, 2), "GB/s": round(gb / t, 1)}), flush=True)
t = timed(both)
print(json.dumps({"what": "H2D + D2H at once", "ms": round(t * 1e3, 2), "GB/s each": round(gb / t, 1)}), flush=True)

solver = MixedPrecisionMultigrid(tolerance=1e-8, device=dev)
solver.setup(n, n)
b64 = solver._engine.levels[0].bufs(torch.float64)
ops.fill_sinsin_(b64.f, (0.0, 1.0, 0.0, 1.0), 2 * 3.141592653589793 ** 2, 1.0, 1.0)
h_in.copy_(b64.f)
prob = PoissonProblem(rhs=h_in, nx=n, ny=n)
solver.solve(prob, reuse_output=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
_, info = solver.solve(prob, reuse_output=True)
t = time.perf_counter() - t0
print(json.dumps({"what": "solve(reuse_output=True)", "ms": round(t * 1e3, 2), "cycles_ms": round(info["cycle_time"] * 1e3, 2),
                  "iterations": info["iterations"]}), flush=True)


def both_under_solve():
    both()
    solver._solve_device(False)


t = timed(both_under_solve)
print(json.dumps({"what": "H2D + D2H at once while a solve cycles", "ms": round(t * 1e3, 2),
                  "GB/s each": round(gb / t, 1)}), flush=True)
d_pitched = b64.tmp  # a pitched device field of the solver (not touched by fresh solves until the first update)
scratch_in = torch.empty_like(d_in)


def under_solve(fn):
    def run():
        fn()
        solver._solve_device(False)
    return run


def ce_h2d():
    with torch.cuda.stream(s1):
        scratch_in.copy_(h_in, non_blocking=True)


def ce_d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


for name, fn in (("CE H2D under solve", under_solve(ce_h2d)), ("CE D2H under solve", under_solve(ce_d2h))):
    t = timed(fn)
    print(json.dumps({"what": name, "ms": round(t * 1e3, 2), "GB/s": round(gb / t, 1)}), flush=True)
# (An SM-driven copy kernel -- 8..128 CTAs, 8 independent 8-byte loads per thread, device <-> pinned host -- was
# measured here too and dropped: 44 GB/s per direction alone, 25 GB/s each when both directions run, no better under
# a solve; profiles/r02_pcie_probe.log keeps those lines.)
for B in (4, 10, 20):
    outs = [h_out, torch.empty((n, n), dtype=torch.float64, pin_memory=True)] if B == 4 else outs
    solver.solve_many([prob] * 2, outputs=outs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    solver.solve_many([prob] * B, outputs=[outs[k % 2] for k in range(B)])
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    print(json.dumps({"what": f"solve_many batch of {B}", "ms_per_solve": round(t * 1e3 / B, 2),
                      "steady_state_ms_estimate": None}), flush=True)
