"""Host-side logic of the multi-GPU path on CPU: world_size-2 (and 4) gloo process groups drive the real
DistributedCycleEngine / DistributedMixedPrecisionSolver with the oracle standing in for the slab kernels.
Acceptance: N-rank result == 1-rank result on every owned row, identical cycle counts."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mixed_precision_multigrid_solvers_for_pdes_b200.distributed import (GHOST, SlabPartition,  # noqa: E402
                                                                         choose_dist_levels)
from oracle import np_oracle as O  # noqa: E402


def test_partition_geometry():
    for world in (1, 2, 4, 8):
        nx, ny, L = world * 1024 + 1, 1025, 9
        D = choose_dist_levels(nx, ny, world, L, agglomerate_below=129)
        assert D >= 2
        owned = [set() for _ in range(D)]
        for r in range(world):
            p = SlabPartition(nx, ny, world, r, L, D, (0, world, 0, 1))
            for l in range(D):
                s = p.slab(l)
                assert s.row0 % 2 == 0 and s.hx == s.hy
                assert s.g_lo == (0 if r == 0 else GHOST) and s.g_hi == (0 if r == world - 1 else GHOST)
                rows = set(range(s.own_lo, s.own_hi))
                assert not (rows & owned[l])
                owned[l] |= rows
                off, n = p.coarse_view(l)
                c = p.slab(l + 1)
                # coarse local row `off + ic` is global coarse row (row0 + 2 ic) / 2
                assert c.row0 + off == s.row0 // 2 and off + n <= c.loc_nx
        for l in range(D):
            assert owned[l] == set(range((nx - 1) // 2 ** l + 1))
    with pytest.raises(ValueError):
        SlabPartition(1003, 1025, 2, 0, 5, 3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dist_emulation import OracleBackend
        from mixed_precision_multigrid_solvers_for_pdes_b200.distributed import (DistributedCycleEngine,
                                                                                 DistributedMixedPrecisionSolver)
        nx, ny, dom = case["nx"], case["ny"], case["domain"]
        g = O.OGrid(nx, ny, dom)
        x, y = g.coords()
        f = 2 * np.pi ** 2 * np.sin(np.pi * x)[:, None] * np.sin(np.pi * y)[None, :]
        transport = None
        if case.get("shm_dir"):
            from mixed_precision_multigrid_solvers_for_pdes_b200.halo import FileShmTransport
            transport = FileShmTransport(case["shm_dir"], GHOST)
        coef = (lambda X, Y: 1.0 + 0.5 * np.sin(2 * np.pi * X) * np.cos(np.pi * Y) + 0.25 * X * Y) if case.get("varcoef") else None
        if case["kind"] == "cycle":
            eng = DistributedCycleEngine(nx, ny, domain=dom, cycle_type=case["cycle"], agglomerate_below=case["agg"],
                                         backend=OracleBackend(), transport=transport, coefficient=coef,
                                         shift=case.get("shift", 0.0))
            b = eng.bufs(0, torch.float64)
            s = eng.part.slab(0)
            b.f.copy_(torch.from_numpy(f[s.row0:s.row0 + s.loc_nx]))
            eng.set_valid(b.f, GHOST)  # filled from the global array, ghost rows included
            ss = torch.zeros(1, dtype=torch.float64)
            norms = []
            for _ in range(case["cycles"]):
                eng.cycle(torch.float64, 0, sumsq_out=ss)
                eng.allreduce_sum(ss)
                norms.append(float(np.sqrt(s.hx * s.hy * ss.item())))
            u = eng.gather_solution(eng.bufs(0, torch.float64).u)
            res = {"norms": norms, "u": u.numpy(), "D": eng.D, "L": eng.num_levels, "ex": eng.exchanges,
                   "pushes": transport.pushes if transport is not None else 0}
        elif case["kind"] == "heat":
            from mixed_precision_multigrid_solvers_for_pdes_b200.distributed import DistributedHeatSolver
            hs = DistributedHeatSolver(tolerance=case["tol"], precision_strategy=case["strategy"],
                                       use_fused_defect_down="always",
                                       agglomerate_below=case["agg"], backend=OracleBackend())
            r = hs.solve_heat_problem(_heat_problem(dom, case.get("source", False)), nx, ny, _time_config(case))
            res = {"u": r["final_solution"], "iters": r["total_mg_iterations"], "steps": r["total_steps"],
                   "errors": r["errors"], "t": r["final_time"], "ex": r["halo_exchanges"], "rows": r["local_rows"],
                   "keys": sorted(r.keys())}
        else:
            sol = DistributedMixedPrecisionSolver(nx, ny, domain=dom, precision_strategy=case["strategy"],
                                                  use_fused_defect_down="always",
                                                  tolerance=1e-8, agglomerate_below=case["agg"], backend=OracleBackend(),
                                                  transport=transport, coefficient=coef)
            sol.set_rhs_from_global(f)
            u, info = sol.solve()
            res = {"norms": info["residual_history"], "u": sol.eng.gather_solution(u).numpy(), "D": sol.eng.D,
                   "L": sol.eng.num_levels, "ex": info["halo_exchanges"], "switches": info["precision_switches"]}
        if rank == 0:
            torch.save(res, out)
    finally:
        dist.destroy_process_group()


def _heat_problem(dom, with_source=False):
    """u = sin(kx x) sin(ky y) exp(-alpha (kx^2+ky^2) t) [+ a steady forced mode], zero on the boundary of `dom`."""
    from mixed_precision_multigrid_solvers_for_pdes_b200 import HeatProblem
    kx, ky, alpha = np.pi / (dom[1] - dom[0]), np.pi / (dom[3] - dom[2]), 0.7
    mode = lambda x, y: np.sin(kx * (x - dom[0])) * np.sin(ky * (y - dom[2]))  # noqa: E731
    k2 = kx ** 2 + ky ** 2
    if not with_source:
        return HeatProblem("decay", mode, None, lambda x, y, t: mode(x, y) * np.exp(-alpha * k2 * t),
                           thermal_diffusivity=alpha, domain=dom)
    # u = mode * (1 + t): u_t - alpha lap u = mode * (1 + alpha k2 (1 + t))
    return HeatProblem("forced", mode, lambda x, y, t: mode(x, y) * (1 + alpha * k2 * (1 + t)),
                       lambda x, y, t: mode(x, y) * (1 + t), thermal_diffusivity=alpha, domain=dom)


def _time_config(case):
    from mixed_precision_multigrid_solvers_for_pdes_b200 import TimeSteppingConfig, TimeSteppingMethod
    return TimeSteppingConfig(TimeSteppingMethod(case["method"]), dt=case["dt"], t_final=case["t_final"])


def _run(world, case, tmp_path):
    out = str(tmp_path / f"res_{world}.pt")
    mp.spawn(_worker, args=(world, _free_port(), case, out), nprocs=world, join=True)
    return torch.load(out, weights_only=False)


@pytest.mark.parametrize("cycle", ["V", "W"])
def test_two_rank_fp64_cycles_equal_single_process_reference(tmp_path, cycle):
    nx, ny = 257, 129  # (nx-1)/2 = 128 rows per rank; square cells on (0,2)x(0,1)
    case = dict(kind="cycle", nx=nx, ny=ny, domain=(0.0, 2.0, 0.0, 1.0), cycle=cycle, agg=33, cycles=3)
    r2 = _run(2, case, tmp_path)
    assert r2["D"] >= 2 and r2["ex"] > 0
    # single-process oracle: the plain reference recursion on the global grid
    g = O.OGrid(nx, ny, case["domain"])
    x, y = g.coords()
    f = 2 * np.pi ** 2 * np.sin(np.pi * x)[:, None] * np.sin(np.pi * y)[None, :]
    s = O.OracleMultigrid(nx, ny, max_levels=r2["L"], cycle_type=cycle, max_iterations=3, tolerance=0.0,
                          domain=case["domain"])
    u, info = s.solve(f)
    assert np.array_equal(r2["u"], u)                       # owned rows bit-identical to the 1-process run
    np.testing.assert_allclose(r2["norms"], info["residual_history"], rtol=1e-13)


@pytest.mark.parametrize("strategy", ["double", "adaptive"])
def test_two_ranks_equal_one_rank_with_variable_coefficients(tmp_path, strategy):
    """-div(a grad u) + shift*u on slabs: coefficient slabs evaluated per level in globally aligned blocks,
    fp64 passes of one sweep (2 passes per smoothing step, each with its own ghost-validity bookkeeping), the full coarse
    coefficient field on the agglomerated levels.  2 ranks == 1 rank bit for bit, same cycle counts."""
    case = dict(kind="solve", nx=257, ny=129, domain=(0.0, 2.0, 0.0, 1.0), agg=33, strategy=strategy, varcoef=True)
    r2, r1 = _run(2, case, tmp_path), _run(1, case, tmp_path)
    assert r2["D"] == r1["D"] >= 2 and r2["ex"] > 0
    assert len(r2["norms"]) == len(r1["norms"]) <= 14 and r1["norms"][-1] < 1e-8
    assert np.array_equal(r2["u"], r1["u"])
    np.testing.assert_allclose(r2["norms"], r1["norms"], rtol=1e-12)
    cyc = dict(kind="cycle", nx=257, ny=129, domain=(0.0, 2.0, 0.0, 1.0), cycle="W", agg=33, cycles=2, varcoef=True, shift=25.0)
    c2, c1 = _run(2, cyc, tmp_path), _run(1, cyc, tmp_path)
    assert np.array_equal(c2["u"], c1["u"])


def test_four_ranks_equal_one_rank(tmp_path):
    case = dict(kind="cycle", nx=513, ny=65, domain=(0.0, 8.0, 0.0, 1.0), cycle="V", agg=17, cycles=2)
    r4 = _run(4, case, tmp_path)
    r1 = _run(1, case, tmp_path)
    assert r4["D"] == r1["D"] >= 2
    assert np.array_equal(r4["u"], r1["u"])
    np.testing.assert_allclose(r4["norms"], r1["norms"], rtol=1e-13)


def test_two_rank_mixed_precision_solve(tmp_path):
    case = dict(kind="solve", nx=257, ny=129, domain=(0.0, 2.0, 0.0, 1.0), strategy="adaptive", agg=33)
    r2 = _run(2, case, tmp_path)
    r1 = _run(1, case, tmp_path)
    assert len(r2["norms"]) == len(r1["norms"]) and r2["norms"][-1] < 1e-8
    np.testing.assert_allclose(r2["norms"], r1["norms"], rtol=1e-9)
    assert np.array_equal(r2["u"], r1["u"])
    assert [s["iteration"] for s in r2["switches"]] == [s["iteration"] for s in r1["switches"]]
    exact = np.sin(np.pi * np.linspace(0, 2, 257))[:, None] * np.sin(np.pi * np.linspace(0, 1, 129))[None, :]
    assert np.max(np.abs(r2["u"] - exact)) < 2e-4


@pytest.mark.parametrize("method,strategy", [("backward_euler", "double"), ("crank_nicolson", "adaptive")])
def test_two_rank_heat_steps_equal_one_rank(tmp_path, method, strategy):
    """BASELINE configs[4] host logic: theta-method stepping on row slabs, one shifted distributed solve per step.
    2 ranks == 1 rank bit for bit (same cycle counts), the last (shortened) step included."""
    case = dict(kind="heat", nx=257, ny=129, domain=(0.0, 2.0, 0.0, 1.0), agg=33, method=method, strategy=strategy,
                dt=0.004, t_final=0.015, tol=1e-9, source=(method == "crank_nicolson"))
    r2 = _run(2, case, tmp_path)
    r1 = _run(1, case, tmp_path)
    assert r2["steps"] == r1["steps"] == 4 and abs(r2["t"] - 0.015) < 1e-14
    assert r2["iters"] == r1["iters"] and r2["ex"] > 0
    assert np.array_equal(r2["u"], r1["u"])
    for k in ("l2_error", "relative_l2_error", "max_error", "relative_max_error"):
        assert abs(r2["errors"][k] - r1["errors"][k]) <= 1e-12 * abs(r1["errors"][k])
    # accuracy: O(dt) resp. O(dt^2) + O(h^2) against the analytical solution
    assert r2["errors"]["relative_max_error"] < (4e-3 if method == "backward_euler" else 2e-4)
    for key in ("problem_name", "grid_size", "time_config", "final_solution", "final_time", "total_steps", "total_time",
                "total_solver_time", "avg_mg_iterations", "total_mg_iterations", "errors", "solver_type", "use_gpu"):
        assert key in r2["keys"], key  # result keys of applications/heat_solver.py:227-247


def test_heat_fp64_step_equals_single_process_oracle(tmp_path):
    """One rank, fp64 strategy: every step is exactly the oracle's shifted multigrid solve started from u^n."""
    dom = (0.0, 2.0, 0.0, 1.0)
    case = dict(kind="heat", nx=129, ny=65, domain=dom, agg=17, method="backward_euler", strategy="double", dt=0.01,
                t_final=0.03, tol=1e-9)
    r = _run(2, case, tmp_path)
    prob = _heat_problem(dom)
    g = O.OGrid(129, 65, dom)
    x, y = g.coords()
    X, Y = np.meshgrid(x, y, indexing="ij")
    u = prob.initial_condition(X, Y)
    u[0, :] = u[-1, :] = 0
    u[:, 0] = u[:, -1] = 0
    lam, iters = 1.0 / (prob.thermal_diffusivity * 0.01), 0
    for _ in range(3):
        rhs = lam * u
        scale = O.l2_norm(rhs, g.hx, g.hy)
        s = O.OracleMultigrid(129, 65, max_levels=5, max_iterations=50, tolerance=1e-9 * scale, domain=dom, shift=lam)
        u, info = s.solve(rhs, initial_guess=u)
        iters += info["iterations"]
    assert r["iters"] == iters
    assert np.max(np.abs(r["u"] - u)) <= 1e-13 * np.max(np.abs(u))


@pytest.mark.parametrize("world", [2, 4])
def test_peer_push_transport_equals_send_recv_exchange(tmp_path, world):
    """halo.py: ghost rows WRITTEN into the neighbours' arrays (ready handshake, push, data handshake) instead of
    send/recv.  File-backed shared mappings stand in for NVLink peer memory; the row bookkeeping and the protocol
    order are the production code.  Result == the send/recv exchange, bit for bit, with the same exchange count."""
    nx = 128 * world + 1
    case = dict(kind="cycle", nx=nx, ny=65, domain=(0.0, float(2 * world), 0.0, 1.0), cycle="V", agg=17, cycles=3)
    ref = _run(world, case, tmp_path)
    shm = tmp_path / "shm"
    shm.mkdir()
    got = _run(world, dict(case, shm_dir=str(shm)), tmp_path)
    assert got["D"] == ref["D"] >= 2 and got["ex"] == ref["ex"] > 0
    assert got["pushes"] >= got["ex"]  # at least one array and one neighbour per exchange on rank 0
    assert np.array_equal(got["u"], ref["u"])
    assert got["norms"] == ref["norms"]


def test_peer_push_transport_mixed_precision_solve(tmp_path):
    case = dict(kind="solve", nx=257, ny=129, domain=(0.0, 2.0, 0.0, 1.0), strategy="adaptive", agg=33)
    ref = _run(2, case, tmp_path)
    shm = tmp_path / "shm"
    shm.mkdir()
    got = _run(2, dict(case, shm_dir=str(shm)), tmp_path)
    assert got["norms"] == ref["norms"] and np.array_equal(got["u"], ref["u"])
