"""NumPy model of the ROW PIPELINE of rbgs_stream_kernel (csrc/mg_stream.cuh), one strip spanning all columns:
sliding windows w / fr, the Jacobi previous-iterate rows pj, stage s on the row of age s, tile lead/tail rows, the
store of the row of age NS, the residual of age NS+1 and the restriction trigger `kpar == (NS & 1)`.
Test infrastructure (tests/test_pipeline_model.py): it proves the schedule against the oracle on the CPU, for
red-black GS (NS = 2 nu) and damped Jacobi (NS = nu, odd NS included), independent of the tile height."""
import numpy as np

from oracle import np_oracle as O

def relax(uc, up, dn, rt, lf, rhs, hx, hy, omega):
    diag = -2.0 / hx ** 2 - 2.0 / hy ** 2
    nb = (up + dn) / hx ** 2 + (rt + lf) / hy ** 2
    un = (rhs + nb) / (-diag)
    return (1 - omega) * uc + omega * un

def model(u, f, hx, hy, omega, nu, jacobi, back, R, RB=4):
    nx, ny = u.shape
    NS = nu if jacobi else 2 * nu
    H = ((NS + (2 if back else 0)) + 1) & ~1
    WR = NS + 2 + (1 if back else 0)
    FR = NS + 1 + (1 if back else 0)
    out = np.full_like(u, np.nan)
    res = np.full_like(u, np.nan)
    nxc, nyc = (nx - 1) // 2 + 1, (ny - 1) // 2 + 1
    rc = np.full((nxc, nyc), np.nan)
    for I0 in range(0, nx, R):
        I1 = min(I0 + R, nx)
        i_begin, i_last = I0 - H, I1 - 1 + H
        nbox = -(-(i_last - i_begin + 1) // RB)
        w = np.zeros((WR, ny)); fr = np.zeros((FR, ny)); pj = np.zeros((NS + 1, ny)); rr = np.zeros((3, ny))
        for i in range(i_begin, i_begin + nbox * RB):
            kpar = (i - i_begin) & 1
            w[1:] = w[:-1].copy(); fr[1:] = fr[:-1].copy()
            w[0] = u[i] if 0 <= i < nx else 0.0
            fr[0] = f[i] if 0 <= i < nx else 0.0
            for s in range(1, NS + 1):
                q = i - s
                if jacobi:
                    old = w[s].copy()
                    if 1 <= q <= nx - 2:
                        new = relax(old[1:-1], w[s - 1][1:-1], pj[s][1:-1], old[2:], old[:-2], fr[s][1:-1], hx, hy, omega)
                        w[s][1:-1] = new
                    pj[s] = old
                elif 1 <= q <= nx - 2:
                    e0 = (kpar + s + ((s - 1) & 1)) & 1
                    J = np.arange(1, ny - 1)
                    J = J[J % 2 == e0]
                    w[s][J] = relax(w[s][J], w[s - 1][J], w[s + 1][J], w[s][J + 1], w[s][J - 1], fr[s][J], hx, hy, omega)
            qf = i - NS
            if I0 <= qf < I1:
                out[qf] = w[NS]
            if back:
                A = NS + 1
                q2 = i - A
                r = np.zeros(ny)
                if 0 <= q2 < nx:
                    r = fr[A].copy()
                    if 1 <= q2 <= nx - 2:
                        c = w[A]
                        lap = (w[A - 1][1:-1] + w[A + 1][1:-1]) / hx ** 2 + (c[2:] + c[:-2]) / hy ** 2 - c[1:-1] * (2 / hx ** 2 + 2 / hy ** 2)
                        r[1:-1] = fr[A][1:-1] - (-1.0) * lap
                if I0 <= q2 < I1:
                    res[q2] = r
                rr[2] = rr[1]; rr[1] = rr[0]; rr[0] = r
                if kpar == (NS & 1):
                    fi = q2 - 1
                    ic = fi >> 1
                    assert fi % 2 == 0
                    if I0 <= fi < I1:
                        v = rr[1][::2].copy()
                        if 0 < ic < nxc - 1:
                            corners = ((rr[2][1:-2:2] + rr[2][3::2]) + rr[0][1:-2:2]) + rr[0][3::2]
                            edges = ((rr[2][2:-1:2] + rr[0][2:-1:2]) + rr[1][1:-2:2]) + rr[1][3::2]
                            v[1:-1] = (0.0625 * corners + 0.125 * edges) + 0.25 * rr[1][2:-1:2]
                        rc[ic] = v
    return out, res, rc

