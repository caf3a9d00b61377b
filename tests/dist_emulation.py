"""CPU stand-in for the fused slab kernels, used ONLY by the gloo tests of the distributed host logic
(partitioning, ghost exchange, agglomeration, norm all-reduce).  It reproduces the SEMANTICS of
mg_vc_pass_slab / mg_vc_defect_pass_slab on dense CPU tensors with the NumPy oracle: the local array is a
stand-alone grid whose first/last rows are never updated, an even row count is allowed for the transfers,
only `norm_rows` enter the residual sum."""
import numpy as np
import torch

from oracle import np_oracle as O


def _np(t):
    return t.numpy()


def _varcoef_cycle(u, f, a, hx, hy, shift, lvl, levels, cycle_type, pre, post, ctol, cmax):
    """The recursion of solvers/multigrid.py:253-337 for -div(a grad u) + shift*u with the oracle's operators
    (coefficients of coarser levels by injection, coarsest level: red-black GS sweeps to tolerance)."""
    if lvl == levels - 1:
        for _ in range(cmax):
            u = O.varcoef_rbgs_smooth(u, f, a, hx, hy, 1.0, 1, shift)
            if O.l2_norm(O.varcoef_residual(u, f, a, hx, hy, shift), hx, hy) < ctol:
                break
        return u
    u = O.varcoef_rbgs_smooth(u, f, a, hx, hy, 1.0, pre, shift)
    rc = O.restrict(O.varcoef_residual(u, f, a, hx, hy, shift))
    ec = np.zeros_like(rc)
    for _ in range({"V": 1, "W": 2}[cycle_type]):
        ec = _varcoef_cycle(ec, rc, np.ascontiguousarray(a[::2, ::2]), 2 * hx, 2 * hy, shift, lvl + 1, levels, cycle_type,
                            pre, post, ctol, cmax)
    u = u + O.prolong(ec)
    return O.varcoef_rbgs_smooth(u, f, a, hx, hy, 1.0, post, shift)


class _CoarseEmu:
    def __init__(self, nx, ny, domain, levels, cycle_type, pre, post, coarse_tol, coarse_max, shift=0.0,
                 coefficient=None):
        self.args = dict(max_levels=levels, cycle_type=cycle_type, pre=pre, post=post, coarse_tolerance=coarse_tol,
                         coarse_max_iterations=coarse_max, domain=domain, shift=shift)
        self.nx, self.ny = nx, ny
        self.coefficient = None if coefficient is None else np.asarray(coefficient, dtype=np.float64)
        self._b = {}

    def bufs(self, dtype):
        if dtype not in self._b:
            class B:
                pass
            b = B()
            b.f = torch.zeros(self.nx, self.ny, dtype=dtype)
            b.u = torch.zeros(self.nx, self.ny, dtype=dtype)
            self._b[dtype] = b
        return self._b[dtype]

    def cycle(self, dtype, u_zero):
        b = self.bufs(dtype)
        npdt = np.float64 if dtype == torch.float64 else np.float32
        if self.coefficient is not None:  # variable coefficients: every level in the level dtype (emulation only)
            A = self.args
            dom = A["domain"]
            hx, hy = (dom[1] - dom[0]) / (self.nx - 1), (dom[3] - dom[2]) / (self.ny - 1)
            u0 = np.zeros_like(_np(b.u)) if u_zero else _np(b.u).copy()
            u = _varcoef_cycle(u0, _np(b.f).copy(), self.coefficient.astype(npdt), npdt(hx), npdt(hy), npdt(A["shift"]), 0,
                               A["max_levels"], A["cycle_type"], A["pre"], A["post"], A["coarse_tolerance"],
                               min(A["coarse_max_iterations"], 200))
            b.u.copy_(torch.from_numpy(np.ascontiguousarray(u)).to(dtype))
            return b.u
        s = O.OracleMultigrid(self.nx, self.ny, dtype=npdt, **self.args)
        L = len(s.grids)
        s.level_dtypes = [npdt] * (L - 1) + [np.float64]
        s.rhs[0] = _np(b.f).copy()
        u0 = np.zeros_like(_np(b.u)) if u_zero else _np(b.u).copy()
        u = s._cycle(u0, 0)
        b.u.copy_(torch.from_numpy(np.ascontiguousarray(u)).to(dtype))
        return b.u


class OracleBackend:
    def empty(self, nx, ny, dtype):
        return torch.zeros(nx, ny, dtype=dtype)

    def scalar(self, n=1):
        return torch.zeros(n, dtype=torch.float64)

    def make_coarse_engine(self, *a, **k):
        return _CoarseEmu(*a, **k)

    def sumsq(self, t):
        return (t.to(torch.float64) ** 2).sum().reshape(1)

    def apply_laplacian(self, u, hx, hy):
        return torch.from_numpy(O.apply_laplacian(_np(u), hx, hy, 1.0))

    def heat_rhs(self, u, rhs, hx, hy, *, lam, c_lap=0.0, f1=None, c_f1=0.0, f0=None, c_f0=0.0, a=None,
                 zero_first_row=True, zero_last_row=True, norm_rows=None):
        U = _np(u)
        t = U.copy()
        if c_lap != 0.0 and a is not None:
            t = t - c_lap * O.varcoef_apply(U, _np(a), hx, hy, 0.0)  # div(a grad u); 0 on the local ring
        elif c_lap != 0.0:
            t = t + c_lap * O.apply_laplacian(U, hx, hy, 1.0)  # 0 on the local first / last rows and columns
        if f1 is not None:
            t = t + c_f1 * _np(f1)
        if f0 is not None:
            t = t + c_f0 * _np(f0)
        t = lam * t
        t[:, 0] = 0
        t[:, -1] = 0
        if zero_first_row:
            t[0, :] = 0
        if zero_last_row:
            t[-1, :] = 0
        rhs.copy_(torch.from_numpy(t))
        lo, hi = norm_rows if norm_rows is not None else (0, t.shape[0])
        return torch.tensor([float(np.sum(t[lo:hi] ** 2))], dtype=torch.float64)

    def vc_pass(self, u_in, u_out, f, hx, hy, *, sweeps=2, omega=1.0, coefficient=-1.0, coarse_in=None,
                coarse_out=None, sumsq_out=None, u_zero=False, norm_rows=None, rows=0, shift=0.0, workspace=None, a=None):
        F = _np(f)
        nx = F.shape[0]
        nxo = nx if nx % 2 == 1 else nx - 1
        U = np.zeros_like(F) if u_zero else _np(u_in).copy()
        if coarse_in is not None:
            C = _np(coarse_in)[:(nxo - 1) // 2 + 1]
            U[:nxo] += O.prolong(C)
        dt = F.dtype.type
        if a is not None:
            U = O.varcoef_rbgs_smooth(U, F, _np(a), dt(hx), dt(hy), dt(omega), sweeps, dt(shift))
        else:
            U = O.rbgs_smooth(U, F, hx, hy, omega, sweeps, shift)
        if u_out is not None:
            u_out.copy_(torch.from_numpy(U))
        if coarse_out is not None or sumsq_out is not None:
            R = (O.varcoef_residual(U, F, _np(a), dt(hx), dt(hy), dt(shift)) if a is not None
                 else O.residual(U, F, hx, hy, coefficient, shift))
            if coarse_out is not None:
                rc = O.restrict(R[:nxo])
                coarse_out[:rc.shape[0]].copy_(torch.from_numpy(rc))
            if sumsq_out is not None:
                lo, hi = norm_rows if norm_rows is not None else (0, nx)
                sumsq_out[0] = float(np.sum(R[lo:hi].astype(np.float64) ** 2))

    def vc_defect_pass(self, u_in, u_out, f, hx, hy, *, e_in=None, r_out=None, sumsq_out=None, coefficient=-1.0,
                       norm_rows=None, rows=0, shift=0.0, workspace=None, u_zero=False, a=None):
        U = np.zeros_like(_np(f)) if u_zero else _np(u_in).copy()
        if e_in is not None:
            U = U + _np(e_in).astype(np.float64)
            u_out.copy_(torch.from_numpy(U))
        if r_out is not None:
            R = (O.varcoef_residual(U, _np(f), _np(a), hx, hy, shift) if a is not None
                 else O.residual(U, _np(f), hx, hy, coefficient, shift))
            r_out.copy_(torch.from_numpy(R.astype(np.float32)))
            lo, hi = norm_rows if norm_rows is not None else (0, U.shape[0])
            sumsq_out[0] = float(np.sum(R[lo:hi] ** 2))

    def vc_defect_down_pass(self, u_in, u_out, f, hx, hy, *, e_in=None, r_out=None, e_out=None, coarse_out=None,
                            sumsq_out=None, omega=1.0, coefficient=-1.0, shift=0.0, u_zero=False, norm_rows=None, rows=0,
                            workspace=None):
        """The fused defect + down pass is, by construction, the defect pass followed by the first pass of the fp32
        error cycle (two sweeps from zero, residual, restriction)."""
        self.vc_defect_pass(u_in, u_out, f, hx, hy, e_in=e_in, r_out=r_out, sumsq_out=sumsq_out, coefficient=coefficient,
                            norm_rows=norm_rows, shift=shift, u_zero=u_zero)
        self.vc_pass(None, e_out, r_out, hx, hy, sweeps=2, omega=omega, coefficient=coefficient, coarse_out=coarse_out,
                     u_zero=True, shift=shift)
