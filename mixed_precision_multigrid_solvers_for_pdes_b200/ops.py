"""Functional layer over the C ABI: one Python function per libmgb200 entry point, operating on
pitched CUDA tensors (see device.py).  Functions ending in ``_`` work in place.  All launches go
to torch's current stream, so they compose with CUDA-graph capture and torch.distributed."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _lib
from .device import code, empty_field, ld, stream_ptr, torch_dtype

_ws: Dict[Tuple[int, int], torch.Tensor] = {}


def _workspace(dev: torch.device) -> torch.Tensor:
    """Per-(device, stream) reduction scratch: partials followed by 8 result slots."""
    # one scratch per device: launches are stream-ordered and the package drives one stream per device
    # (a capturing stream replays in the same order), so the buffer is never used concurrently
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), 0)
    w = _ws.get(key)
    if w is None:
        n = _lib.call("mg_sumsq_workspace_doubles")
        w = torch.zeros(n + 8, dtype=torch.float64, device=dev)
        _ws[key] = w
    return w


def _check_same_shape(a: torch.Tensor, b: torch.Tensor, what: str) -> None:
    if a.shape != b.shape:
        raise ValueError(f"{what}: shapes {tuple(a.shape)} and {tuple(b.shape)} differ")


def apply_laplacian(u: torch.Tensor, hx: float, hy: float, coefficient: float = 1.0,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    nx, ny = u.shape
    if out is None:
        out = empty_field(nx, ny, u.dtype, u.device, zero=False)
    _lib.call("mg_apply_laplacian", u.data_ptr(), out.data_ptr(), nx, ny, ld(u), ld(out), hx, hy, coefficient,
              code(u.dtype), stream_ptr())
    return out


def residual(u: torch.Tensor, f: torch.Tensor, hx: float, hy: float, coefficient: float = 1.0,
             out: Optional[torch.Tensor] = None, out_dtype=None, shift: float = 0.0) -> torch.Tensor:
    _check_same_shape(u, f, "residual")
    if u.dtype != f.dtype:
        raise TypeError("residual: u and f must share a dtype")
    nx, ny = u.shape
    if out is None:
        out = empty_field(nx, ny, out_dtype or u.dtype, u.device, zero=False)
    _lib.call("mg_residual_h", u.data_ptr(), f.data_ptr(), out.data_ptr(), nx, ny, ld(u), ld(f), ld(out), hx, hy,
              coefficient, shift, code(u.dtype), code(out.dtype), stream_ptr())
    return out


def smooth_rbgs_(u: torch.Tensor, f: torch.Tensor, hx: float, hy: float, omega: float = 1.0, sweeps: int = 1,
                 shift: float = 0.0):
    _check_same_shape(u, f, "smooth_rbgs")
    nx, ny = u.shape
    _lib.call("mg_smooth_rbgs_h", u.data_ptr(), f.data_ptr(), nx, ny, ld(u), ld(f), hx, hy, omega, shift, sweeps,
              code(u.dtype), stream_ptr())
    return u


def smooth_jacobi_(u: torch.Tensor, f: torch.Tensor, hx: float, hy: float, omega: float = 2.0 / 3.0,
                   sweeps: int = 1, tmp: Optional[torch.Tensor] = None):
    _check_same_shape(u, f, "smooth_jacobi")
    nx, ny = u.shape
    if tmp is None or ld(tmp) != ld(u):
        tmp = torch.empty_strided((nx, ny), (ld(u), 1), dtype=u.dtype, device=u.device)
    _lib.call("mg_smooth_jacobi", u.data_ptr(), tmp.data_ptr(), f.data_ptr(), nx, ny, ld(u), ld(f), hx, hy, omega,
              sweeps, code(u.dtype), stream_ptr())
    return u


LEXGS_MODES = {"forward": 0, "backward": 1, "symmetric": 2}


def smooth_lexgs_(u: torch.Tensor, f: torch.Tensor, hx: float, hy: float, omega: float = 1.0, sweeps: int = 1,
                  mode: str = "forward"):
    _check_same_shape(u, f, "smooth_lexgs")
    nx, ny = u.shape
    _lib.call("mg_smooth_lexgs", u.data_ptr(), f.data_ptr(), nx, ny, ld(u), ld(f), hx, hy, omega, sweeps,
              LEXGS_MODES[mode], code(u.dtype), stream_ptr())
    return u


def coarse_solve_lexgs_(u: torch.Tensor, f: torch.Tensor, hx: float, hy: float, omega: float = 1.0,
                        coefficient: float = -1.0, tolerance: float = 1e-12, max_iterations: int = 1000,
                        info: Optional[torch.Tensor] = None, shift: float = 0.0):
    """`info`: optional device tensor of >= 2 float64 receiving (sweeps done, last norm)."""
    _check_same_shape(u, f, "coarse_solve")
    nx, ny = u.shape
    _lib.call("mg_coarse_solve_lexgs_h", u.data_ptr(), f.data_ptr(), nx, ny, ld(u), ld(f), hx, hy, omega, coefficient,
              shift, tolerance, max_iterations, info.data_ptr() if info is not None else None, code(u.dtype), stream_ptr())
    return u


def restrict(fine: torch.Tensor, method: str = "full_weighting", out: Optional[torch.Tensor] = None,
             out_dtype=None) -> torch.Tensor:
    nxf, nyf = fine.shape
    nxc, nyc = (nxf - 1) // 2 + 1, (nyf - 1) // 2 + 1
    if out is None:
        out = empty_field(nxc, nyc, out_dtype or fine.dtype, fine.device, zero=False)
    elif tuple(out.shape) != (nxc, nyc):
        raise ValueError(f"Cannot restrict from {tuple(fine.shape)} to {tuple(out.shape)}")
    _lib.call("mg_restrict", fine.data_ptr(), out.data_ptr(), nxf, nyf, ld(fine), ld(out), _lib.RESTRICT[method],
              code(fine.dtype), code(out.dtype), stream_ptr())
    return out


def prolong(coarse: torch.Tensor, method: str = "bilinear", out: Optional[torch.Tensor] = None, add: bool = False,
            out_dtype=None) -> torch.Tensor:
    nxc, nyc = coarse.shape
    nxf, nyf = 2 * (nxc - 1) + 1, 2 * (nyc - 1) + 1
    if out is None:
        if add:
            raise ValueError("prolong(add=True) needs the fine field to add to")
        out = empty_field(nxf, nyf, out_dtype or coarse.dtype, coarse.device, zero=False)
    elif tuple(out.shape) != (nxf, nyf):
        raise ValueError(f"Cannot prolongate from {tuple(coarse.shape)} to {tuple(out.shape)}")
    _lib.call("mg_prolong", coarse.data_ptr(), out.data_ptr(), nxc, nyc, ld(coarse), ld(out), _lib.PROLONG[method],
              1 if add else 0, code(coarse.dtype), code(out.dtype), stream_ptr())
    return out


def sumsq_async(x: torch.Tensor, slot: int = 0) -> torch.Tensor:
    """Launch the deterministic sum of squares; returns a 1-element device view (no sync)."""
    w = _workspace(x.device)
    n = w.numel() - 8
    out = w[n + slot:n + slot + 1]
    nx, ny = x.shape
    _lib.call("mg_sumsq", x.data_ptr(), nx, ny, ld(x), code(x.dtype), w.data_ptr(), out.data_ptr(), stream_ptr())
    return out


_host_slots: Dict[int, torch.Tensor] = {}


def read_scalar(x: torch.Tensor) -> float:
    """Value of a 1-element float64 device tensor on the host.  The value travels by a kernel store into page-locked
    host memory (mg_read_doubles), not by the copy engine, so it cannot queue behind a bulk download that another
    stream has in flight (MixedPrecisionMultigrid.solve_many); only the CURRENT stream is synchronised."""
    dev = x.device.index if x.device.index is not None else torch.cuda.current_device()
    h = _host_slots.get(dev)
    if h is None:
        h = _host_slots[dev] = torch.zeros(8, dtype=torch.float64, pin_memory=True)
    _lib.call("mg_read_doubles", x.data_ptr(), h.data_ptr(), 1, stream_ptr())
    torch.cuda.current_stream(x.device).synchronize()
    return float(h[0])


def sumsq(x: torch.Tensor) -> float:
    return float(sumsq_async(x).item())


def cast(x: torch.Tensor, dtype, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    tdt = torch_dtype(dtype)
    if out is None:
        if x.dtype == tdt:
            return x
        out = empty_field(x.shape[0], x.shape[1], tdt, x.device, zero=False)
    nx, ny = x.shape
    _lib.call("mg_cast", x.data_ptr(), out.data_ptr(), nx, ny, ld(x), ld(out), code(x.dtype), code(out.dtype),
              stream_ptr())
    return out


def axpy_(alpha: float, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    _check_same_shape(x, y, "axpy")
    nx, ny = x.shape
    _lib.call("mg_axpy", float(alpha), x.data_ptr(), y.data_ptr(), nx, ny, ld(x), ld(y), code(x.dtype),
              code(y.dtype), stream_ptr())
    return y


def zero_(x: torch.Tensor) -> torch.Tensor:
    """x = 0 over the whole pitched extent of the field (mg_zero: one memset node, no fill kernel)."""
    nx = x.shape[0]
    _lib.call("mg_zero", x.data_ptr(), nx, ld(x), code(x.dtype), stream_ptr())
    return x


def zero_ring_(x: torch.Tensor, first_row: bool = True, last_row: bool = True) -> torch.Tensor:
    """Zero the first / last column and (optionally: a row slab only owns them on the physical boundary) the first /
    last row of a field, in one launch."""
    nx, ny = x.shape
    _lib.call("mg_zero_ring", x.data_ptr(), nx, ny, ld(x), 1 if first_row else 0, 1 if last_row else 0, code(x.dtype),
              stream_ptr())
    return x


def heat_rhs_(u: torch.Tensor, rhs: torch.Tensor, hx: float, hy: float, *, lam: float, c_lap: float = 0.0,
              f1: Optional[torch.Tensor] = None, c_f1: float = 0.0, f0: Optional[torch.Tensor] = None, c_f0: float = 0.0,
              a: Optional[torch.Tensor] = None, zero_first_row: bool = True, zero_last_row: bool = True,
              norm_rows: Optional[Tuple[int, int]] = None, slot: int = 3) -> torch.Tensor:
    """rhs = lam * (u + c_lap * L_h u + c_f1 * f1 + c_f0 * f0) with the boundary ring zeroed, in one pass (mg_heat_rhs);
    returns a 1-element device view holding sum(rhs^2) over `norm_rows` (no sync)."""
    for t in (u, rhs, f1, f0, a):
        if t is not None and t.dtype != torch.float64:
            raise TypeError("heat_rhs_: fp64 fields")
    nx, ny = u.shape
    w = _workspace(u.device)
    n = w.numel() - 8
    out = w[n + slot:n + slot + 1]
    lo, hi = norm_rows if norm_rows is not None else (0, -1)
    _lib.call("mg_heat_rhs", u.data_ptr(), f1.data_ptr() if f1 is not None else None,
              f0.data_ptr() if f0 is not None else None, a.data_ptr() if a is not None else None, rhs.data_ptr(),
              out.data_ptr(), w.data_ptr(), nx, ny, ld(u), ld(f1) if f1 is not None else 0, ld(f0) if f0 is not None else 0,
              ld(a) if a is not None else 0, ld(rhs), hx, hy, float(c_lap), float(c_f1), float(c_f0), float(lam),
              1 if zero_first_row else 0, 1 if zero_last_row else 0, lo, hi, stream_ptr())
    return out


def fill_sinsin_(f: torch.Tensor, domain=(0.0, 1.0, 0.0, 1.0), amplitude: float = 1.0, kx: float = 1.0,
                 ky: float = 1.0) -> torch.Tensor:
    nx, ny = f.shape
    _lib.call("mg_fill_sinsin", f.data_ptr(), nx, ny, ld(f), float(domain[0]), float(domain[1]), float(domain[2]),
              float(domain[3]), float(amplitude), float(kx), float(ky), code(f.dtype), stream_ptr())
    return f


def maxerr_sinsin(u: torch.Tensor, domain=(0.0, 1.0, 0.0, 1.0), amplitude: float = 1.0, kx: float = 1.0,
                  ky: float = 1.0) -> float:
    w = _workspace(u.device)
    n = w.numel() - 8
    out = w[n + 7:n + 8]
    nx, ny = u.shape
    _lib.call("mg_maxerr_sinsin", u.data_ptr(), nx, ny, ld(u), float(domain[0]), float(domain[1]), float(domain[2]),
              float(domain[3]), float(amplitude), float(kx), float(ky), code(u.dtype), w.data_ptr(), out.data_ptr(),
              stream_ptr())
    return float(out.item())


# ------------------------------------------------------------------------------------------------
# Fused / temporally blocked passes (mg_vc_* family)
# ------------------------------------------------------------------------------------------------
_vc_ws: Dict[Tuple[int, int], torch.Tensor] = {}
_vc_ws_retired: list = []  # outgrown scratch buffers: never freed, a captured CUDA graph may still write to them


def vc_workspace_doubles(nx: int, ny: int) -> int:
    return _lib.call("mg_vc_workspace_doubles", nx, ny)


def _vc_workspace(dev: torch.device, nx: int, ny: int) -> torch.Tensor:
    """Default reduction scratch of the fused passes for callers that do not bring their own (`workspace=`):
    one per (device, stream), grow-only.  A buffer that has been outgrown is retired, not freed -- CUDA graphs
    captured earlier have its address baked into their norm passes.  Solver objects own their scratch
    (CycleEngine.workspace), so two solvers never share partials."""
    need = vc_workspace_doubles(nx, ny)
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(dev).cuda_stream)
    w = _vc_ws.get(key)
    if w is None or w.numel() < need:
        if w is not None:
            _vc_ws_retired.append(w)
        w = torch.zeros(need, dtype=torch.float64, device=dev)
        _vc_ws[key] = w
    return w


def vc_aligned(*fields) -> bool:
    """True when every field satisfies the vector path's 16-byte base / pitch alignment."""
    for t in fields:
        if t is None:
            continue
        va = 16 // t.element_size()
        if t.data_ptr() % 16 or ld(t) % va:
            return False
    return True


class KernelTimer:
    """Optional CUDA-event timing of individual fused passes inside a running solve (bench.py's roofline
    leg): events are recorded on the launching stream around every vc_pass on a field of at least
    `min_points` points; `summary()` synchronises and returns per-kernel launch counts and mean ms."""

    def __init__(self, min_points: int = 0):
        self.min_points = min_points
        self.records = []  # (tag, start_event, end_event)

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for tag, a, b in self.records:
            d = out.setdefault(tag, {"launches": 0, "total_ms": 0.0})
            d["launches"] += 1
            d["total_ms"] += a.elapsed_time(b)
        for d in out.values():
            d["mean_ms"] = d["total_ms"] / d["launches"]
        return out


TIMER: Optional[KernelTimer] = None


def vc_pass(u_in: torch.Tensor, u_out: Optional[torch.Tensor], f: torch.Tensor, hx: float, hy: float, *,
            sweeps: int = 2, omega: float = 1.0, coefficient: float = -1.0,
            coarse_in: Optional[torch.Tensor] = None, coarse_out: Optional[torch.Tensor] = None,
            sumsq_out: Optional[torch.Tensor] = None, loader: str = "tma", rows: int = 0,
            u_zero: bool = False, norm_rows: Optional[Tuple[int, int]] = None, shift: float = 0.0,
            smoother: str = "rbgs", workspace: Optional[torch.Tensor] = None, a: Optional[torch.Tensor] = None) -> None:
    """One fused pass: [u += P coarse_in] -> `sweeps` RB-GS (``smoother="jacobi"``: damped Jacobi) sweeps ->
    [coarse_out = R(f - A u)] or [sumsq_out[0] = sum (f - A u)^2]; out of place u_in -> u_out (u_out=None:
    nothing stored).  ``u_zero``: treat u_in as identically zero without reading it.  ``a``: nodal coefficient field
    of the level -> the variable-coefficient operator -div(a grad u) + shift*u (mg_vcv_pass_slab; RB-GS only, fp64 at
    most one sweep per pass; `coefficient` is not used)."""
    nx, ny = f.shape
    flags = (rows & 0xFFF) << 8
    if smoother == "jacobi":
        flags |= _lib.VC_JACOBI
    elif smoother != "rbgs":
        raise ValueError(f"vc_pass: smoother must be 'rbgs' or 'jacobi', got {smoother!r}")
    if coarse_in is not None:
        flags |= _lib.VC_PROLONG
    if coarse_out is not None:
        flags |= _lib.VC_RESTRICT
    if sumsq_out is not None:
        flags |= _lib.VC_NORM
    if u_out is None:
        flags |= _lib.VC_NO_STORE
    if u_zero:
        flags |= _lib.VC_U_ZERO
    flags |= _loader_flag(loader)
    ws = None
    if sumsq_out is not None:
        ws = workspace if workspace is not None else _vc_workspace(f.device, nx, ny)
        if ws.numel() < vc_workspace_doubles(nx, ny):
            raise ValueError("vc_pass: workspace too small for this grid (see vc_workspace_doubles)")
    timed = TIMER is not None and nx * ny >= TIMER.min_points
    if timed:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    nlo, nhi = norm_rows if norm_rows is not None else (0, -1)
    if a is not None:
        if a.shape != f.shape or a.dtype != f.dtype:
            raise ValueError("vc_pass: the coefficient field must have the shape and dtype of the level")
        _lib.call("mg_vcv_pass_slab", None if u_zero else u_in.data_ptr(), u_out.data_ptr() if u_out is not None else None,
                  f.data_ptr(), a.data_ptr(),
                  coarse_in.data_ptr() if coarse_in is not None else None,
                  coarse_out.data_ptr() if coarse_out is not None else None,
                  sumsq_out.data_ptr() if sumsq_out is not None else None, ws.data_ptr() if ws is not None else None,
                  nx, ny, 0 if u_zero else ld(u_in), ld(u_out) if u_out is not None else 0, ld(f), ld(a),
                  ld(coarse_in) if coarse_in is not None else 0, ld(coarse_out) if coarse_out is not None else 0,
                  hx, hy, omega, sweeps, code(f.dtype), flags, nlo, nhi, shift, stream_ptr())
    else:
        _vc_pass_const(u_in, u_out, f, hx, hy, omega, coefficient, coarse_in, coarse_out, sumsq_out, ws, sweeps, flags,
                       nlo, nhi, shift, u_zero)
    if timed:
        ev1.record()
        tag = (("var:" if a is not None else "") + ("Z+" if u_zero else "") + ("P+" if coarse_in is not None else "")
               + f"{'jac' if smoother == 'jacobi' else 'rbgs'}{sweeps}"
               + ("+R" if coarse_out is not None else "") + ("+N" if sumsq_out is not None else "")
               + f"/{'f64' if f.dtype == torch.float64 else 'f32'}/{nx}x{ny}")
        TIMER.records.append((tag, ev0, ev1))


def _vc_pass_const(u_in, u_out, f, hx, hy, omega, coefficient, coarse_in, coarse_out, sumsq_out, ws, sweeps, flags, nlo, nhi,
                   shift, u_zero) -> None:
    nx, ny = f.shape
    _lib.call("mg_vc_pass_slab", None if u_zero else u_in.data_ptr(), u_out.data_ptr() if u_out is not None else None,
              f.data_ptr(),
              coarse_in.data_ptr() if coarse_in is not None else None,
              coarse_out.data_ptr() if coarse_out is not None else None,
              sumsq_out.data_ptr() if sumsq_out is not None else None, ws.data_ptr() if ws is not None else None,
              nx, ny, 0 if u_zero else ld(u_in), ld(u_out) if u_out is not None else 0, ld(f),
              ld(coarse_in) if coarse_in is not None else 0, ld(coarse_out) if coarse_out is not None else 0,
              hx, hy, omega, coefficient, sweeps, code(f.dtype), flags, nlo, nhi, shift, stream_ptr())


def _loader_flag(loader: str) -> int:
    if loader == "cp_async":
        return _lib.VC_LOADER_CPASYNC
    if loader != "tma":
        raise ValueError(f"unknown loader {loader!r}")
    return 0


def vc_defect_pass(u_in: torch.Tensor, u_out: Optional[torch.Tensor], f: torch.Tensor, hx: float, hy: float, *,
                   e_in: Optional[torch.Tensor] = None, r_out: Optional[torch.Tensor] = None,
                   sumsq_out: Optional[torch.Tensor] = None, coefficient: float = -1.0, loader: str = "tma",
                   rows: int = 0, norm_rows: Optional[Tuple[int, int]] = None, shift: float = 0.0,
                   workspace: Optional[torch.Tensor] = None, u_zero: bool = False,
                   a: Optional[torch.Tensor] = None) -> None:
    """Mixed-precision defect-correction pass on the fp64 iterate (one HBM pass):
    u_out = u_in + e_in (fp32 correction; None: u unchanged, nothing stored), r_out = fp32(f - A u_out),
    sumsq_out[0] = sum of the squared fp64 residual."""
    nx, ny = u_in.shape
    if u_in.dtype != torch.float64 or f.dtype != torch.float64:
        raise TypeError("vc_defect_pass: the iterate and right-hand side are fp64")
    for t in (e_in, r_out):
        if t is not None and t.dtype != torch.float32:
            raise TypeError("vc_defect_pass: correction and residual are fp32")
    flags = ((rows & 0xFFF) << 8) | _loader_flag(loader)
    ws = None
    if r_out is not None:
        ws = workspace if workspace is not None else _vc_workspace(u_in.device, nx, ny)
        if ws.numel() < vc_workspace_doubles(nx, ny):
            raise ValueError("vc_defect_pass: workspace too small for this grid (see vc_workspace_doubles)")
    if u_zero:
        flags |= _lib.VC_U_ZERO
    timed = TIMER is not None and nx * ny >= TIMER.min_points
    if timed:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    nlo, nhi = norm_rows if norm_rows is not None else (0, -1)
    if a is not None:  # variable coefficients: the fp64 nodal field of the level
        if a.shape != u_in.shape or a.dtype != torch.float64:
            raise ValueError("vc_defect_pass: the coefficient field is fp64 and has the shape of the iterate")
        _lib.call("mg_vcv_defect_pass_slab", u_in.data_ptr(), u_out.data_ptr() if u_out is not None else None, f.data_ptr(),
                  a.data_ptr(), e_in.data_ptr() if e_in is not None else None,
                  r_out.data_ptr() if r_out is not None else None,
                  sumsq_out.data_ptr() if sumsq_out is not None else None, ws.data_ptr() if ws is not None else None,
                  nx, ny, ld(u_in), ld(u_out) if u_out is not None else 0, ld(f), ld(a),
                  ld(e_in) if e_in is not None else 0, ld(r_out) if r_out is not None else 0, hx, hy, flags, nlo, nhi,
                  shift, stream_ptr())
    else:
        _lib.call("mg_vc_defect_pass_slab", u_in.data_ptr(), u_out.data_ptr() if u_out is not None else None, f.data_ptr(),
                  e_in.data_ptr() if e_in is not None else None, r_out.data_ptr() if r_out is not None else None,
                  sumsq_out.data_ptr() if sumsq_out is not None else None, ws.data_ptr() if ws is not None else None,
                  nx, ny, ld(u_in), ld(u_out) if u_out is not None else 0, ld(f), ld(e_in) if e_in is not None else 0,
                  ld(r_out) if r_out is not None else 0, hx, hy, coefficient, flags, nlo, nhi, shift, stream_ptr())
    if timed:
        ev1.record()
        TIMER.records.append((("var:" if a is not None else "") + ("Z+" if u_zero else "") + ("update+" if e_in is not None else "")
                              + ("resid32+N" if r_out is not None else "") + f"/f64/{nx}x{ny}", ev0, ev1))


def vc_defect_down_pass(u_in: torch.Tensor, u_out: Optional[torch.Tensor], f: torch.Tensor, hx: float, hy: float, *,
                        e_in: Optional[torch.Tensor], r_out: torch.Tensor, e_out: torch.Tensor, coarse_out: torch.Tensor,
                        sumsq_out: torch.Tensor, omega: float = 1.0, coefficient: float = -1.0, shift: float = 0.0,
                        u_zero: bool = False, norm_rows: Optional[Tuple[int, int]] = None, rows: int = 0,
                        workspace: Optional[torch.Tensor] = None) -> None:
    """Defect pass + pre-smoothing pass of the error equation in one HBM pass (mg_vc_defect_down_pass_slab):
    u_out = u_in + e_in (e_in None: u unchanged, nothing stored); r_out = fp32(f - A u_out); sumsq_out[0] = sum r^2 (fp64);
    e_out = 2 RB-GS sweeps from zero on A e = r_out; coarse_out = R(r_out - A e_out).  Bit-identical to vc_defect_pass
    followed by vc_pass(u_zero=True, sweeps=2, coarse_out=...)."""
    nx, ny = f.shape
    if f.dtype != torch.float64 or (not u_zero and u_in.dtype != torch.float64):
        raise TypeError("vc_defect_down_pass: the iterate and right-hand side are fp64")
    for t in (e_in, r_out, e_out, coarse_out):
        if t is not None and t.dtype != torch.float32:
            raise TypeError("vc_defect_down_pass: correction, residual, error iterate and coarse right-hand side are fp32")
    flags = ((rows & 0xFFF) << 8) | (_lib.VC_U_ZERO if u_zero else 0)
    ws = workspace if workspace is not None else _vc_workspace(f.device, nx, ny)
    if ws.numel() < vc_workspace_doubles(nx, ny):
        raise ValueError("vc_defect_down_pass: workspace too small for this grid (see vc_workspace_doubles)")
    timed = TIMER is not None and nx * ny >= TIMER.min_points
    if timed:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    nlo, nhi = norm_rows if norm_rows is not None else (0, -1)
    _lib.call("mg_vc_defect_down_pass_slab", None if u_zero else u_in.data_ptr(),
              u_out.data_ptr() if (u_out is not None and e_in is not None) else None, f.data_ptr(),
              e_in.data_ptr() if e_in is not None else None, r_out.data_ptr(), e_out.data_ptr(), coarse_out.data_ptr(),
              sumsq_out.data_ptr(), ws.data_ptr(), nx, ny, 0 if u_zero else ld(u_in),
              ld(u_out) if (u_out is not None and e_in is not None) else 0, ld(f), ld(e_in) if e_in is not None else 0,
              ld(r_out), ld(e_out), ld(coarse_out), hx, hy, omega, coefficient, flags, nlo, nhi, shift, stream_ptr())
    if timed:
        ev1.record()
        TIMER.records.append(("dd:" + ("Z+" if u_zero else "") + ("update+" if e_in is not None else "")
                              + f"resid32+N+rbgs2+R/f64/{nx}x{ny}", ev0, ev1))


def varcoef_coarse_solve_(u: torch.Tensor, f: torch.Tensor, a: torch.Tensor, hx: float, hy: float, shift: float = 0.0,
                          omega: float = 1.0, tolerance: float = 1e-12, max_iterations: int = 1000,
                          info: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Coarsest-level solve for -div(a grad u) + shift*u: red-black GS sweeps to `tolerance` in one launch."""
    _check_same_shape(u, f, "varcoef_coarse_solve")
    nx, ny = u.shape
    _lib.call("mg_varcoef_coarse_solve", u.data_ptr(), f.data_ptr(), a.data_ptr(), nx, ny, ld(u), ld(f), ld(a), hx, hy,
              shift, omega, tolerance, max_iterations, info.data_ptr() if info is not None else None, code(u.dtype),
              stream_ptr())
    return u


SMALL_CYCLE_SMEM_LIMIT = 200 * 1024
_CYCLES = {"V": 0, "W": 1, "F": 2}


def small_cycle_fits(nx: int, ny: int, nlev: int, dtype, coarse_dtype) -> bool:
    if nlev < 1 or nlev > 8:
        return False
    b = _lib.call("mg_small_cycle_smem_bytes", nx, ny, nlev, code(dtype), code(coarse_dtype))
    return 0 < b <= SMALL_CYCLE_SMEM_LIMIT


def small_cycle_(u: torch.Tensor, f: torch.Tensor, hx: float, hy: float, *, nlev: int, cycle_type: str = "V",
                 pre: int = 2, post: int = 2, omega: float = 1.0, coefficient: float = -1.0, shift: float = 0.0,
                 coarse_tolerance: float = 1e-12, coarse_max_iterations: int = 1000, coarse_dtype=None,
                 u_zero: bool = False, info: Optional[torch.Tensor] = None, profile: bool = False) -> torch.Tensor:
    """One complete sub-cycle over `nlev` levels in a single launch (mg_small_cycle), in place on u.
    ``profile``: `info` (16 doubles) accumulates SM cycles per phase kind (see include/mgb200.h)."""
    if profile and (info is None or info.numel() < 16):
        raise ValueError("small_cycle_(profile=True) needs an info tensor of 16 doubles")
    nx, ny = u.shape
    cd = coarse_dtype if coarse_dtype is not None else u.dtype
    timed = TIMER is not None and nx * ny >= TIMER.min_points
    if timed:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    _lib.call("mg_small_cycle", u.data_ptr(), f.data_ptr(), nx, ny, ld(u), ld(f), hx, hy, nlev, _CYCLES[cycle_type],
              pre, post, omega, coefficient, shift, coarse_tolerance, coarse_max_iterations,
              (1 if u_zero else 0) | (2 if profile else 0),
              info.data_ptr() if info is not None else None, code(u.dtype), code(cd), stream_ptr())
    if timed:
        ev1.record()
        TIMER.records.append((f"small{cycle_type}{nlev}/{'f64' if u.dtype == torch.float64 else 'f32'}/{nx}x{ny}", ev0, ev1))
    return u
