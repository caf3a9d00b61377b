"""Multi-GPU path on real GPUs.  The single-process tests run the distributed engine with world = 1 (slab code path,
agglomeration, coarse views) against the single-GPU engine; the 2-rank test is launched with torchrun when at least two
GPUs are visible (`gpurun --gpus 2`)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_world1_distributed_engine_equals_oracle():
    from mixed_precision_multigrid_solvers_for_pdes_b200.distributed import DistributedMixedPrecisionSolver
    n = 513
    f = O.mms_rhs(n)
    for strategy, cycles in (("double", 8), ("adaptive", 8)):
        sol = DistributedMixedPrecisionSolver(n, n, precision_strategy=strategy, tolerance=1e-8, agglomerate_below=65,
                                              device=torch.device("cuda", 0), use_fused_defect_down="always")
        assert sol.eng.D >= 2 and sol.fused_defect_down == (strategy == "adaptive")
        sol.set_rhs_from_global(torch.from_numpy(f).cuda())
        u, info = sol.solve()
        assert info["converged"] and info["iterations"] == cycles
        if strategy == "double":
            ou, oinfo = O.OracleMultigrid(n, max_levels=sol.eng.num_levels).solve(f)
            np.testing.assert_allclose(info["residual_history"], oinfo["residual_history"], rtol=1e-11)
            assert np.max(np.abs(u.cpu().numpy() - ou)) <= 1e-12
        err = np.max(np.abs(u.cpu().numpy() - O.mms_exact(n)))
        assert abs(err - O.mms_discretisation_error(n)) <= 0.01 * O.mms_discretisation_error(n)


def test_world1_distributed_heat_equals_single_gpu_heat_solver():
    """The slab heat driver (world = 1) against HeatSolver2D on the same problem: same steps, same accuracy."""
    from mixed_precision_multigrid_solvers_for_pdes_b200 import (HeatProblem, HeatSolver2D, TimeSteppingConfig,
                                                                 TimeSteppingMethod)
    from mixed_precision_multigrid_solvers_for_pdes_b200.distributed import DistributedHeatSolver
    n, alpha = 257, 0.5
    mode = lambda X, Y: np.sin(np.pi * X) * np.sin(np.pi * Y)  # noqa: E731
    exact = lambda X, Y, t: mode(X, Y) * np.exp(-2 * np.pi ** 2 * alpha * t)  # noqa: E731
    prob = HeatProblem("pure_diffusion", mode, None, exact, thermal_diffusivity=alpha)
    for method, bound in ((TimeSteppingMethod.BACKWARD_EULER, 1e-2), (TimeSteppingMethod.CRANK_NICOLSON, 2e-4)):
        cfg = TimeSteppingConfig(method, dt=0.005, t_final=0.02)
        rd = DistributedHeatSolver(tolerance=1e-10, agglomerate_below=33, device=torch.device("cuda", 0),
                                   use_cuda_graphs=True, use_fused_defect_down="always").solve_heat_problem(prob, n, n, cfg)
        rs = HeatSolver2D(tolerance=1e-10).solve_heat_problem(prob, n, n, cfg)
        assert rd["total_steps"] == rs["total_steps"] == 4
        assert np.max(np.abs(rd["final_solution"] - rs["final_solution"])) <= 1e-8
        assert rd["errors"]["relative_max_error"] < bound
        assert abs(rd["errors"]["max_error"] - rs["errors"]["max_error"]) <= 1e-6 * rs["errors"]["max_error"] + 1e-9


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpus_equal_one_gpu(tmp_path):
    out = str(tmp_path / "res.pt")
    script = os.path.join(ROOT, "tests", "dist_gpu_worker.py")
    for world in (1, 2):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), script, out + str(world)]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    r1 = torch.load(out + "1", weights_only=False)
    r2 = torch.load(out + "2", weights_only=False)
    assert r2["graphs_captured"] > 0
    for r in (r1, r2):  # CUDA-graph replay (incl. the NCCL exchanges) == eager, bit for bit
        assert np.array_equal(r["adaptive"]["u"], r["adaptive_graphs"]["u"])
        assert r["adaptive"]["history"] == r["adaptive_graphs"]["history"]
    for key in ("double", "adaptive"):
        assert r1[key]["iterations"] == r2[key]["iterations"]
        np.testing.assert_allclose(r2[key]["history"], r1[key]["history"], rtol=1e-10)
        assert np.array_equal(r1[key]["u"], r2[key]["u"]), key   # owned rows bit-identical across GPU counts
    assert r2["exchanges"] > 0
    # heat stepping on slabs (configs[4]): same steps, same cycle counts, bit-identical field, O(dt) accurate
    h1, h2 = r1["heat"], r2["heat"]
    assert h1["steps"] == h2["steps"] == 4 and h1["iters"] == h2["iters"] and h2["exchanges"] > 0
    assert np.array_equal(h1["u"], h2["u"])
    assert h2["errors"]["relative_max_error"] < 2e-3


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpus_peer_push_halo_equals_one_gpu(tmp_path):
    """One-sided ghost-row pushes over NVLink peer memory (halo.py) instead of NCCL send/recv: first run on hardware in
    round 2 (2 x B200, profiles/r02_dist_2gpu_pytest.log), bit-identical to the single-GPU run; strict since."""
    out = str(tmp_path / "res.pt")
    script = os.path.join(ROOT, "tests", "dist_gpu_worker.py")
    for world, halo in ((1, "nccl"), (2, "p2p")):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(29520 + world), script, out + str(world)]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, MGB200_TEST_HALO=halo))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    r1 = torch.load(out + "1", weights_only=False)
    r2 = torch.load(out + "2", weights_only=False)
    assert r2["pushes"] > 0 and r2["graphs_captured"] > 0
    for key in ("double", "adaptive", "adaptive_graphs"):
        assert r1[key]["iterations"] == r2[key]["iterations"]
        assert np.array_equal(r1[key]["u"], r2[key]["u"]), key


def test_world1_distributed_variable_coefficients_match_the_single_gpu_facade():
    """-div(a grad u) on row slabs (world = 1: slab passes, coefficient slabs per level, agglomerated coarse engine with the
    full coarse field) against MixedPrecisionMultigrid(coefficient=...): same cycle counts, same solution to rounding
    (the facade injects the coarse coefficients from its fine-grid evaluation, the slabs evaluate every level)."""
    from mixed_precision_multigrid_solvers_for_pdes_b200 import MixedPrecisionMultigrid, PoissonProblem
    from mixed_precision_multigrid_solvers_for_pdes_b200.distributed import DistributedMixedPrecisionSolver
    n = 513
    coef = lambda X, Y: 1.0 + 0.5 * np.sin(2 * np.pi * X) * np.cos(np.pi * Y) + X * Y  # noqa: E731
    f = O.mms_rhs(n)
    f[0, :] = f[-1, :] = f[:, 0] = f[:, -1] = 0.0
    for strategy in ("double", "adaptive"):
        sol = DistributedMixedPrecisionSolver(n, n, precision_strategy=strategy, tolerance=1e-8, agglomerate_below=65,
                                              device=torch.device("cuda", 0), coefficient=coef)
        sol.set_rhs_from_global(torch.from_numpy(f).cuda())
        u, info = sol.solve()
        us, si = MixedPrecisionMultigrid(strategy, coefficient=coef, tolerance=1e-8).solve(PoissonProblem(rhs=f, nx=n, ny=n))
        assert info["converged"] and info["iterations"] == si["iterations"], (info["iterations"], si["iterations"])
        np.testing.assert_allclose(info["residual_history"], si["residual_history"], rtol=1e-6)
        assert np.max(np.abs(u.cpu().numpy() - us)) <= 1e-11 * np.max(np.abs(us))


def test_world1_solve_many_on_slabs_is_solve_per_right_hand_side():
    """The pipelined batch API of the slab solver (pinned host slabs in, pinned host slabs out, transfers overlapping the
    cycles): every solution equals the one `solve()` returns for that right-hand side, bit for bit."""
    from mixed_precision_multigrid_solvers_for_pdes_b200.distributed import DistributedMixedPrecisionSolver
    n = 513
    sol = DistributedMixedPrecisionSolver(n, n, precision_strategy="adaptive", tolerance=1e-8, agglomerate_below=65,
                                          device=torch.device("cuda", 0), use_cuda_graphs=True,
                                          use_fused_defect_down="always")
    base = torch.from_numpy(O.mms_rhs(n))
    fs = [(base * s).pin_memory() for s in (1.0, -2.5, 0.5, 3.0)]
    us = [torch.empty((n, n), dtype=torch.float64).pin_memory() for _ in fs]
    infos = sol.solve_many(fs, us)
    torch.cuda.synchronize()
    assert len(infos) == 4 and all(i["converged"] for i in infos)
    for f, u in zip(fs, us):
        sol.set_rhs_from_global(f.cuda())
        ref, info = sol.solve()
        assert torch.equal(ref.cpu(), u)
    err = np.max(np.abs(us[0].numpy() - O.mms_exact(n)))
    assert abs(err - O.mms_discretisation_error(n)) <= 0.01 * O.mms_discretisation_error(n)
    with pytest.raises(ValueError):
        sol.solve_many(fs, us[:2])
