"""Micro-benchmark of the fused passes (CUDA events, L2-exceeding inputs). Not the contract bench."""
import argparse
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mixed_precision_multigrid_solvers_for_pdes_b200 import ops  # noqa: E402
from mixed_precision_multigrid_solvers_for_pdes_b200.device import empty_field  # noqa: E402


ITERS, WARM = 5, 2


def timeit(fn, iters=None, warm=None):
    iters = ITERS if iters is None else iters
    warm = WARM if warm is None else warm
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16385)
    ap.add_argument("--dtypes", default="f32,f64")
    ap.add_argument("--loaders", default="tma,cp_async")
    ap.add_argument("--rows", default="0")
    ap.add_argument("--modes", default="smooth2,smooth1,down2,up2,up2norm,resrestrict")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--warm", type=int, default=2)
    a = ap.parse_args()
    global ITERS, WARM
    ITERS, WARM = a.iters, a.warm
    n = a.n
    h = 1.0 / (n - 1)
    nc = (n - 1) // 2 + 1
    for dn in a.dtypes.split(","):
        dt = torch.float32 if dn == "f32" else torch.float64
        w = 4 if dn == "f32" else 8
        gen = torch.Generator(device="cuda").manual_seed(0)
        u, f, out = (empty_field(n, n, dt) for _ in range(3))
        u.copy_(torch.rand((n, n), generator=gen, device="cuda", dtype=dt) * 2 - 1)
        f.copy_(torch.rand((n, n), generator=gen, device="cuda", dtype=dt) * 2 - 1)
        ec, rc = empty_field(nc, nc, dt), empty_field(nc, nc, dt)
        ec.copy_(torch.rand((nc, nc), generator=gen, device="cuda", dtype=dt))
        ss = torch.zeros(1, dtype=torch.float64, device="cuda")
        # reference point: plain device copy
        med, best = timeit(lambda: out.copy_(u))
        print(json.dumps({"kernel": "torch_copy", "dtype": dn, "n": n, "ms": round(med, 4),
                          "GBs": round(2 * n * n * w / med / 1e6, 1)}), flush=True)
        for loader in a.loaders.split(","):
            for rows in [int(r) for r in a.rows.split(",")]:
                for mode in a.modes.split(","):
                    kw = dict(loader=loader, rows=rows)
                    if mode == "smooth2":
                        fn, byt, sw = (lambda: ops.vc_pass(u, out, f, h, h, sweeps=2, **kw)), 3 * w, 2
                    elif mode == "smooth1":
                        fn, byt, sw = (lambda: ops.vc_pass(u, out, f, h, h, sweeps=1, **kw)), 3 * w, 1
                    elif mode == "down2":
                        fn, byt, sw = (lambda: ops.vc_pass(u, out, f, h, h, sweeps=2, coarse_out=rc, **kw)), 3.25 * w, 2
                    elif mode == "up2":
                        fn, byt, sw = (lambda: ops.vc_pass(u, out, f, h, h, sweeps=2, coarse_in=ec, **kw)), 3.25 * w, 2
                    elif mode == "up2norm":
                        fn, byt, sw = (lambda: ops.vc_pass(u, out, f, h, h, sweeps=2, coarse_in=ec, sumsq_out=ss, **kw)), 3.25 * w, 2
                    elif mode == "zdown2":
                        fn, byt, sw = (lambda: ops.vc_pass(u, out, f, h, h, sweeps=2, coarse_out=rc, u_zero=True, **kw)), 2.25 * w, 2
                    elif mode == "defect" and dn == "f64":
                        e32 = empty_field(n, n, torch.float32); r32 = empty_field(n, n, torch.float32)
                        fn = (lambda: ops.vc_defect_pass(u, out, f, h, h, e_in=e32, r_out=r32, sumsq_out=ss, **kw))
                        byt, sw = 32.0, 0
                    elif mode.startswith("jac") and loader == "tma":  # damped Jacobi (omega = 2/3) in the streaming kernel
                        jk = dict(omega=2.0 / 3.0, smoother="jacobi", **kw)
                        if mode == "jac2":
                            fn, byt, sw = (lambda: ops.vc_pass(u, out, f, h, h, sweeps=2, **jk)), 3 * w, 2
                        elif mode == "jacdown2":
                            fn, byt, sw = (lambda: ops.vc_pass(u, out, f, h, h, sweeps=2, coarse_out=rc, **jk)), 3.25 * w, 2
                        elif mode == "jacup2":
                            fn, byt, sw = (lambda: ops.vc_pass(u, out, f, h, h, sweeps=2, coarse_in=ec, **jk)), 3.25 * w, 2
                        else:
                            continue
                    elif mode.startswith("v") and mode[1:] in ("smooth1", "smooth2", "zdown2", "zdown1", "up2", "up1", "defect") \
                            and loader == "tma":
                        # variable coefficients -div(a grad u): the nodal field rides along (+1 word per point)
                        if "acoef" not in locals() or acoef.dtype != dt:
                            acoef = empty_field(n, n, dt)
                            acoef.copy_(1.0 + torch.rand((n, n), generator=gen, device="cuda", dtype=dt))
                        m = mode[1:]
                        if m == "defect":
                            if dn != "f64":
                                continue
                            e32 = empty_field(n, n, torch.float32); r32 = empty_field(n, n, torch.float32)
                            fn = (lambda: ops.vc_defect_pass(u, out, f, h, h, e_in=e32, r_out=r32, sumsq_out=ss, a=acoef, rows=rows))
                            byt, sw = 40.0, 0
                        else:
                            sw = int(m[-1])
                            if dn == "f64" and sw > 1:
                                continue
                            if m.startswith("smooth"):
                                fn, byt = (lambda: ops.vc_pass(u, out, f, h, h, sweeps=sw, a=acoef, rows=rows)), 4 * w
                            elif m.startswith("zdown"):
                                fn, byt = (lambda: ops.vc_pass(u, out, f, h, h, sweeps=sw, coarse_out=rc, u_zero=True, a=acoef,
                                                               rows=rows)), 3.25 * w
                            else:
                                fn, byt = (lambda: ops.vc_pass(u, out, f, h, h, sweeps=sw, coarse_in=ec, a=acoef, rows=rows)), 4.25 * w
                    elif mode == "resrestrict":
                        fn, byt, sw = (lambda: ops.vc_pass(u, None, f, h, h, sweeps=0, coarse_out=rc, **kw)), 2.25 * w, 0
                    else:
                        continue
                    med, best = timeit(fn)
                    print(json.dumps({"kernel": mode, "dtype": dn, "loader": loader, "rows": rows, "n": n,
                                      "ms": round(med, 4), "best_ms": round(best, 4),
                                      "hbm_GBs": round(byt * n * n / med / 1e6, 1),
                                      "alg_smoother_GBs": round(3 * w * sw * n * n / med / 1e6, 1)}), flush=True)
        # baseline: basic (unfused) kernels, 2 sweeps in place
        med, best = timeit(lambda: ops.smooth_rbgs_(out, f, h, h, 1.0, 2))
        print(json.dumps({"kernel": "basic_rbgs2", "dtype": dn, "n": n, "ms": round(med, 4),
                          "alg_smoother_GBs": round(3 * w * 2 * n * n / med / 1e6, 1)}), flush=True)
        if any(m.startswith("jac") for m in a.modes.split(",")):
            tmp = empty_field(n, n, dt)
            med, best = timeit(lambda: ops.smooth_jacobi_(out, f, h, h, 2.0 / 3.0, 2, tmp=tmp))
            print(json.dumps({"kernel": "basic_jacobi2", "dtype": dn, "n": n, "ms": round(med, 4),
                              "alg_smoother_GBs": round(3 * w * 2 * n * n / med / 1e6, 1)}), flush=True)
            del tmp
        del u, f, out, ec, rc
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
