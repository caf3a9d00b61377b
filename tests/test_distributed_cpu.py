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
        if case["kind"] == "cycle":
            eng = DistributedCycleEngine(nx, ny, domain=dom, cycle_type=case["cycle"], agglomerate_below=case["agg"],
                                         backend=OracleBackend())
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
            res = {"norms": norms, "u": u.numpy(), "D": eng.D, "L": eng.num_levels, "ex": eng.exchanges}
        else:
            sol = DistributedMixedPrecisionSolver(nx, ny, domain=dom, precision_strategy=case["strategy"],
                                                  tolerance=1e-8, agglomerate_below=case["agg"], backend=OracleBackend())
            sol.set_rhs_from_global(f)
            u, info = sol.solve()
            res = {"norms": info["residual_history"], "u": sol.eng.gather_solution(u).numpy(), "D": sol.eng.D,
                   "L": sol.eng.num_levels, "ex": info["halo_exchanges"], "switches": info["precision_switches"]}
        if rank == 0:
            torch.save(res, out)
    finally:
        dist.destroy_process_group()


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
