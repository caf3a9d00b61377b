"""Per-level device time of one mixed-precision refinement cycle (CUDA events around every pass, eager launches).
Passes shorter than ~10 us are dominated by launch latency when timed this way; the ncu launch list
(profiles/r02_launches_bench.md) is the reference for those.

    python tools/level_profile.py [n]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mixed_precision_multigrid_solvers_for_pdes_b200 import MixedPrecisionMultigrid, ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16385
s = MixedPrecisionMultigrid(tolerance=1e-8, switch_threshold=1e-6)
s.setup(n, n)
b64 = s._engine.levels[0].bufs(torch.float64)
ops.fill_sinsin_(b64.f, (0.0, 1.0, 0.0, 1.0), 2 * np.pi ** 2, 1.0, 1.0)
ops.zero_ring_(b64.f)
s._refinement_residual(u_zero=True)
for k in range(4):
    s._cycle_refinement(u_zero=(k == 0))
lt = s.profile_levels(cycles=5)
pts = [(l.grid.nx, l.grid.ny) for l in s._engine.levels]
tot = sum(v["smooth_time"] for v in lt.values())
for lvl, v in lt.items():
    nx, ny = pts[lvl]
    print(json.dumps({"level": lvl, "grid": [nx, ny], "passes": v["passes"], "ms": round(v["smooth_time"] * 1e3, 4),
                      "share": round(v["smooth_time"] / tot, 4),
                      "ns_per_point": round(v["smooth_time"] * 1e9 / (nx * ny), 4)}))
print(json.dumps({"total_ms": round(tot * 1e3, 4)}))
