"""NumPy replay of the schedule of `lexgs_pipe_kernel` (csrc/mg_lexgs.cuh): warps of 32 rows, lane l at column t - l in
step t, 8 prefetched values per lane and sub-block, the neighbour rows fetched once per 32 steps.  Warps are replayed
one after the other, which is a valid order of the kernel's execution (dependencies only point to lower warp
indices), so a schedule error shows up as a difference from the sequential sweep.  Test infrastructure only."""
import numpy as np


def relax(c, a, b, rt, lf, rhs, hx2, hy2, nd, omega):
    nb = (a + b) / hx2 + (rt + lf) / hy2
    return (1 - omega) * c + omega * ((rhs + nb) / nd)


def sweep(u, f, hx, hy, omega, forward=True):
    u = u.copy()
    nx, ny = u.shape
    nrows, ncols = nx - 2, ny - 2
    hx2, hy2 = hx ** 2, hy ** 2
    nd = 2 / hx2 + 2 / hy2
    I = (lambda lr: 1 + lr) if forward else (lambda lr: nx - 2 - lr)   # sweep-order row -> storage row
    J = (lambda lc: 1 + lc) if forward else (lambda lc: ny - 2 - lc)
    for w in range((nrows + 31) // 32):
        last_lr = min(w * 32 + 31, nrows - 1)
        last_lane = last_lr - w * 32
        rows = [I(w * 32 + l if w * 32 + l < nrows else last_lr) for l in range(32)]
        ok = [w * 32 + l < nrows for l in range(32)]
        uprev, unext = I(w * 32 - 1), I(last_lr + 1)
        nsteps = ncols + last_lane
        prev = [u[rows[l], J(-1)] for l in range(32)]
        carry = [u[rows[l], J(0)] for l in range(32)]

        def load8(tc):
            un = [[0.0] * 8 for _ in range(32)]
            fc = [[0.0] * 8 for _ in range(32)]
            for l in range(32):
                for k in range(8):
                    c = tc + k - l
                    if 0 <= c + 1 <= ncols:
                        un[l][k] = u[rows[l], J(c + 1)]
                    if 0 <= c < ncols:
                        fc[l][k] = f[rows[l], J(c)]
            return un, fc

        nxt = load8(0)
        for tb in range(0, nsteps, 32):
            up32 = [u[uprev, J(tb + l)] if tb + l < ncols else 0.0 for l in range(32)]
            dn32 = [u[unext, J(tb + l - last_lane)] if 0 <= tb + l - last_lane < ncols else 0.0 for l in range(32)]
            for sub in range(4):
                tc = tb + 8 * sub
                if tc >= nsteps:
                    break
                un, fc = nxt
                nxt = load8(tc + 8)  # issued before the 8 steps, like the kernel's double buffer
                for k in range(8):
                    before = [prev[l - 1] if l > 0 else up32[8 * sub + k] for l in range(32)]
                    after = [un[l + 1][k] if l < 31 else 0.0 for l in range(32)]
                    after[last_lane] = dn32[8 * sub + k]
                    for l in range(32):
                        c = tc + k - l
                        if ok[l] and 0 <= c < ncols:
                            res = relax(carry[l], after[l], before[l], un[l][k], prev[l], fc[l][k], hx2, hy2, nd, omega)
                            u[rows[l], J(c)] = res
                            prev[l] = res
                            carry[l] = un[l][k]
    return u
