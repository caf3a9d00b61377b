"""NumPy replay of the row schedule of `defect_down_kernel` (csrc/mg_stream_dd.cuh), one tile of rows at a time:
rows I0-6 .. I1-1+6 (+8 in the last tile) stream through; when row i arrives
    u_new(i) = u(i) + e(i);  r32(i-1) = fp32(f - A u_new)(i-1);  e'(i-1) := 0
    half-sweep stage s = 1..4 relaxes the points of its colour in row (i-1) - s of the error equation A e' = r32
    row (i-1) - 4 of e' is final;  the residual of the error equation is taken on row (i-1) - 5
    and the full weighting centred on row (i-1) - 6 is emitted when that row is even.
Everything a tile has not produced itself is ZERO in its windows (nothing is loaded there), which is what makes the
width of the row halo a correctness question.  Columns are not tiled here (the strips' column halo follows the same
cone).  Test infrastructure only."""
import numpy as np

P = 16  # padding rows of the tile-local windows: window row P + i holds global row i


def _residual_row(W, rhs_row, i, nx, hx, hy, shift, dt):
    """Row i of rhs - A w (A = -lap_h + shift) from the padded window W, r = rhs on the boundary ring; dtype dt."""
    r = rhs_row.astype(dt).copy()
    if 1 <= i <= nx - 2:
        hx2, hy2 = dt(hx) ** 2, dt(hy) ** 2
        c = W[P + i, 1:-1]
        lap = (W[P + i + 1, 1:-1] + W[P + i - 1, 1:-1]) / hx2 + (W[P + i, 2:] + W[P + i, :-2]) / hy2 \
            - c * (dt(2.0) / hx2 + dt(2.0) / hy2)
        au = -lap
        if shift:
            au = au + dt(shift) * c
        r[1:-1] = rhs_row[1:-1].astype(dt) - au
    return r


def _relax_row(E, R, q, colour, hx, hy, shift):
    """Points with (q + j) % 2 == colour of row q of the fp32 error iterate (padded windows), in place, omega = 1."""
    ny = E.shape[1]
    dt = np.float32
    hx2, hy2 = dt(hx) ** 2, dt(hy) ** 2
    diag = dt(2.0) / hx2 + dt(2.0) / hy2 + dt(shift)
    j0 = 1 if (q + 1) % 2 == colour else 2
    J, Jr, Jl = slice(j0, ny - 1, 2), slice(j0 + 1, ny, 2), slice(j0 - 1, ny - 2, 2)
    if E[P + q, J].size:
        nb = (E[P + q + 1, J] + E[P + q - 1, J]) / hx2 + (E[P + q, Jr] + E[P + q, Jl]) / hy2
        E[P + q, J] = (R[P + q, J] + nb) / diag


def fused_defect_down(u, f, e, hx, hy, rows, shift=0.0, lead=6, tail=6, tail_last=8):
    """Returns (u_new, r32, sum r64^2 over all rows, e', f_c) assembled from tiles of `rows` rows."""
    nx, ny = f.shape
    nxc, nyc = (nx + 1) // 2, (ny + 1) // 2
    u_out = np.full((nx, ny), np.nan)
    r_out = np.full((nx, ny), np.nan, dtype=np.float32)
    e_out = np.full((nx, ny), np.nan, dtype=np.float32)
    c_out = np.full((nxc, nyc), np.nan, dtype=np.float32)
    sumsq = 0.0
    for I0 in range(0, nx, rows):
        I1 = min(I0 + rows, nx)
        i_begin, i_last = I0 - lead, I1 - 1 + (tail_last if I1 % 2 else tail)
        Un = np.zeros((nx + 2 * P, ny))                       # u_new rows this tile has formed
        R32 = np.zeros((nx + 2 * P, ny), dtype=np.float32)    # fp32 residual rows (right-hand side of the error equation)
        Ep = np.zeros((nx + 2 * P, ny), dtype=np.float32)     # error iterate
        RR = np.zeros((nx + 2 * P, ny), dtype=np.float32)     # residual rows of the error equation
        for i in range(i_begin, i_last + 1):
            if 0 <= i < nx:
                Un[P + i] = u[i] + e[i].astype(np.float64)
                if I0 <= i < I1:
                    u_out[i] = Un[P + i]
            ip = i - 1
            if 0 <= ip < nx:
                r64 = _residual_row(Un, f[ip], ip, nx, hx, hy, shift, np.float64)
                R32[P + ip] = r64.astype(np.float32)
                if I0 <= ip < I1:
                    r_out[ip] = R32[P + ip]
                    sumsq += float(np.sum(r64 ** 2))
            for s in range(1, 5):
                q = ip - s
                if 1 <= q <= nx - 2:
                    _relax_row(Ep, R32, q, (s - 1) % 2, hx, hy, shift)
            qf = ip - 4
            if 0 <= qf < nx and I0 <= qf < I1:
                e_out[qf] = Ep[P + qf]
            q2 = ip - 5
            if -P < q2 < nx + P - 1:
                if 0 <= q2 < nx:  # rows outside the grid stay zero (the kernel's TMA zero fill)
                    RR[P + q2] = _residual_row(Ep, R32[P + q2], q2, nx, hx, hy, shift, np.float32)
                fi = q2 - 1
                if q2 % 2 == 1 and I0 <= fi < I1:
                    ic = fi // 2
                    a, b, c = RR[P + fi - 1], RR[P + fi], RR[P + fi + 1]
                    row = b.copy()  # boundary injection on the ring of the coarse grid
                    if 0 < ic < nxc - 1:
                        J = np.arange(2, ny - 1, 2)
                        corners = ((a[J - 1] + a[J + 1]) + c[J - 1]) + c[J + 1]
                        edges = ((a[J] + c[J]) + b[J - 1]) + b[J + 1]
                        row[J] = (np.float32(0.0625) * corners + np.float32(0.125) * edges) + np.float32(0.25) * b[J]
                    c_out[ic] = row[::2]
    return u_out, r_out, sumsq, e_out, c_out
