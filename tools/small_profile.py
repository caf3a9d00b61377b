#!/usr/bin/env python
"""Where the time of mg_small_cycle goes: SM-cycle counters per phase kind (profiling flag of the kernel) and CUDA-event
time per launch, for the sub-cycles the solvers actually launch (fp64 from 65^2, fp32 levels + fp64 coarsest from 129^2).

    python tools/small_profile.py > profiles/r02_small_cycle_profile.json
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mixed_precision_multigrid_solvers_for_pdes_b200 import ops  # noqa: E402
from mixed_precision_multigrid_solvers_for_pdes_b200.device import to_device  # noqa: E402

NAMES = ["smooth", "residual+restrict", "prolong", "coarsest_solve"]


def run(n, nlev, dt, cycle, reps=50):
    rng = np.random.default_rng(n)
    f = rng.uniform(-1, 1, (n, n)).astype(dt)
    f[0, :] = f[-1, :] = f[:, 0] = f[:, -1] = 0
    du, df = to_device(np.zeros((n, n), dt))[0], to_device(f)[0]
    h = 1.0 / (n - 1)
    info = torch.zeros(16, dtype=torch.float64, device="cuda")
    kw = dict(nlev=nlev, cycle_type=cycle, coarse_dtype=np.float64, u_zero=True, info=info)
    for _ in range(5):
        ops.small_cycle_(du, df, h, h, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.small_cycle_(du, df, h, h, **kw)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    info.zero_()
    for _ in range(reps):
        ops.small_cycle_(du, df, h, h, profile=True, **kw)
    torch.cuda.synchronize()
    v = info.cpu().numpy()
    launches = v[13]
    total = (v[10] + v[11] + v[12]) / launches
    out = {"entry": f"{n}x{n}", "levels": nlev, "dtype": np.dtype(dt).name, "cycle": cycle, "us_per_launch_events": us,
           "cycles_per_launch": total, "implied_mhz": total / us if us > 0 else None,
           "load": v[10] / launches, "store": v[12] / launches, "last_coarse_sweeps": v[0]}
    for k, name in enumerate(NAMES):
        out["block:" + name] = v[2 + k] / launches
        out["warp0:" + name] = v[6 + k] / launches
    return out


def main():
    rows = []
    for n, nlev, dt in ((65, 5, np.float64), (129, 6, np.float32), (33, 4, np.float64), (17, 3, np.float64)):
        for cycle in ("V", "W"):
            rows.append(run(n, nlev, dt, cycle))
    print(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
