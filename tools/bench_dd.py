"""Times the fused defect + down pass (ops.vc_defect_down_pass) against its two-launch equivalent on an n x n grid
(CUDA events, inputs larger than L2).  Tuning aid for csrc/mg_stream_dd.cu, not the contract bench.

    python tools/bench_dd.py [n] [iters]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mixed_precision_multigrid_solvers_for_pdes_b200 import ops  # noqa: E402
from mixed_precision_multigrid_solvers_for_pdes_b200.device import empty_field  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16385
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
h = 1.0 / (n - 1)
nc = (n + 1) // 2
gen = torch.Generator(device="cuda").manual_seed(0)
u, uo, f = (empty_field(n, n, torch.float64) for _ in range(3))
e, r, eo, tmp = (empty_field(n, n, torch.float32) for _ in range(4))
co = empty_field(nc, nc, torch.float32)
for t in (u, f, e):
    t.copy_(torch.rand((n, n), generator=gen, device="cuda", dtype=t.dtype) * 2 - 1)
    ops.zero_ring_(t)
ss = torch.zeros(1, dtype=torch.float64, device="cuda")


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


ROWS = [int(x) for x in os.environ.get("DD_ROWS", "0").split(",")]
pts = n * n
for rows in ROWS:
    def fused():
        ops.vc_defect_down_pass(u, uo, f, h, h, e_in=e, r_out=r, e_out=eo, coarse_out=co, sumsq_out=ss, rows=rows)

    def defect():
        ops.vc_defect_pass(u, uo, f, h, h, e_in=e, r_out=r, sumsq_out=ss, rows=rows)

    def down():
        ops.vc_pass(tmp, eo, r, h, h, sweeps=2, coarse_out=co, u_zero=True, rows=rows)

    def up():
        ops.vc_pass(eo, tmp, r, h, h, sweeps=2, coarse_in=co, rows=rows)

    out = {"n": n, "rows": rows, "variant": os.environ.get("MG_DD_VARIANT", "0"), "fused_ms": timed(fused)}
    if os.environ.get("MG_DD_VARIANT", "0") == "0":
        out["defect_ms"], out["down_ms"], out["up_ms"] = timed(defect), timed(down), timed(up)
        out["two_launch_ms"] = out["defect_ms"] + out["down_ms"]
    out["fused_gbs"] = 37.0 * pts / out["fused_ms"] / 1e6
    print(json.dumps({k: round(v, 4) if isinstance(v, float) else v for k, v in out.items()}), flush=True)
