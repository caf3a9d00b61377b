"""MixedPrecisionMultigrid facade (README.md:73-92) on the GPU: fp64 strategy == reference runs; mixed
strategies reach the reference tolerance with the same cycle count and an MMS discretisation error
within 1% of the fp64 value (BASELINE.json north star)."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu

from mixed_precision_multigrid_solvers_for_pdes_b200 import (MixedPrecisionMultigrid, PoissonProblem,  # noqa: E402
                                                             PoissonTestProblems)


def source_term(x, y):
    return 2 * np.pi ** 2 * np.sin(np.pi * x) * np.sin(np.pi * y)


def test_readme_example_double_matches_reference(solve_golden, golden_meta):
    problem = PoissonProblem(source_term, nx=129, ny=129)
    solver = MixedPrecisionMultigrid(precision_strategy="double", use_gpu=True)
    solution, info = solver.solve(problem)
    m = [x for x in golden_meta["solves"] if x["name"] == "v129"][0]
    assert info["iterations"] == m["iterations"] == 8 and info["converged"]
    np.testing.assert_allclose(info["residual_history"], solve_golden["v129_hist"], rtol=1e-12)
    assert np.max(np.abs(solution - solve_golden["v129_u"])) <= 1e-12
    for k in ("iterations", "residual", "solve_time", "final_residual", "residual_history", "num_levels"):
        assert k in info
    assert info["residual"] == info["final_residual"] and info["num_levels"] == 6


@pytest.mark.parametrize("n", [129, 1025])
@pytest.mark.parametrize("strategy", ["adaptive", "conservative", "aggressive", "refinement"])
def test_mixed_strategies_converge_like_fp64(n, strategy):
    problem = PoissonProblem(source_term, nx=n, ny=n)
    solver = MixedPrecisionMultigrid(precision_strategy=strategy, switch_threshold=None if strategy != "adaptive" else 1e-6)
    u, info = solver.solve(problem)
    assert info["converged"] and info["final_residual"] < 1e-8
    assert info["iterations"] == 8          # same V-cycle count as the fp64 reference
    assert u.dtype == np.float64
    err = np.max(np.abs(u - O.mms_exact(n)))
    ref = O.mms_discretisation_error(n)      # fp64 reference value (5.0201e-5 at 129^2, 7.8437e-7 at 1025^2)
    assert abs(err - ref) <= 0.01 * ref, (err, ref)
    if strategy == "refinement":
        assert info["precision_switches"] == [] and set(info["precision_history"]) == {"mixed"}
    else:
        sw = info["precision_switches"]
        assert len(sw) == 1 and sw[0]["to"] == "float64" and sw[0]["reason"] == "switch_threshold"
        thr = 1e-4 if strategy == "aggressive" else 1e-6
        assert sw[0]["residual"] <= thr
        k = sw[0]["iteration"]
        assert info["precision_history"][:k] == ["mixed"] * k and set(info["precision_history"][k:]) == {"float64"}
        assert info["residual_history"][k - 2] > thr if k >= 2 else True


def test_single_precision_floors_like_reference():
    # an all-fp32 iterate floors around 1e-3..1e-4 (h-scaled) and never reaches 1e-8
    u, info = MixedPrecisionMultigrid("single", max_iterations=12).solve(PoissonProblem(source_term, nx=129, ny=129))
    assert not info["converged"] and 1e-5 < info["final_residual"] < 1e-2
    assert abs(np.max(np.abs(u - O.mms_exact(129))) - 5.02e-5) < 5e-6


def test_device_generated_rhs_and_w_cycle():
    p = PoissonProblem.manufactured(257, on_device=True)
    u, info = MixedPrecisionMultigrid("adaptive", cycle_type="W").solve(p)
    # fp64 W(2,2) needs 3-4 cycles; the fp32 inner cycle caps the per-cycle reduction near
    # fp32 resolution x stencil amplification, so the refinement phase may take a cycle or two more
    assert info["converged"] and info["iterations"] <= 6
    _, info64 = MixedPrecisionMultigrid("double", cycle_type="W").solve(p)
    assert info64["converged"] and info64["iterations"] <= 4
    assert abs(np.max(np.abs(u - O.mms_exact(257))) - O.mms_discretisation_error(257)) < 1e-9


def test_catalogue_problem_and_errors():
    pr = PoissonTestProblems().get_problem("polynomial")
    u, info = MixedPrecisionMultigrid("adaptive").solve(pr, nx=257, ny=257)
    x = np.linspace(0, 1, 257)
    X, Y = np.meshgrid(x, x, indexing="ij")
    assert info["converged"] and np.max(np.abs(u - pr.analytical_solution(X, Y))) < 2e-6
    # the reference's all-points norm can never converge when f != 0 on the boundary ring
    _, strict = MixedPrecisionMultigrid("double", strict_reference_norm=True, max_iterations=12).solve(pr, nx=257, ny=257)
    assert not strict["converged"] and abs(strict["final_residual"] - 0.0099602384) < 1e-9
    with pytest.raises(ValueError, match="Unknown precision strategy"):
        MixedPrecisionMultigrid("half")
    with pytest.raises(ValueError, match="no CPU path"):
        MixedPrecisionMultigrid(use_gpu=False)
    with pytest.raises(ValueError, match="grid size unknown"):
        MixedPrecisionMultigrid().solve(pr)


def test_large_grid_mixed_reaches_discretisation_accuracy():
    """4097^2: target error 4.9e-8 is BELOW fp32 resolution; the fp64-iterate refinement must reach it."""
    n = 4097
    p = PoissonProblem.manufactured(n, on_device=True)
    s = MixedPrecisionMultigrid("adaptive", tolerance=1e-8)
    p.rhs_array = None
    u, info = s.solve(p)
    assert info["converged"] and info["iterations"] == 8
    from mixed_precision_multigrid_solvers_for_pdes_b200 import ops
    from mixed_precision_multigrid_solvers_for_pdes_b200.device import to_device
    err = ops.maxerr_sinsin(to_device(u)[0])
    # the algebraic error left at ||r|| < 1e-8 is ~1e-10: within 1% of 4.9023e-8
    assert abs(err - O.mms_discretisation_error(n)) <= 0.01 * O.mms_discretisation_error(n)


@pytest.mark.parametrize("strategy", ["adaptive", "double", "refinement"])
def test_cuda_graph_replay_is_bitwise_identical_to_eager(strategy):
    p = PoissonProblem(source_term, nx=513, ny=513)
    ue, ie = MixedPrecisionMultigrid(strategy, use_cuda_graphs=False).solve(p)
    sg = MixedPrecisionMultigrid(strategy, use_cuda_graphs=True)
    ug, ig = sg.solve(p)
    ug = ug.copy()
    ug2, ig2 = sg.solve(p)  # second solve: every step replays a captured graph
    assert ie["residual_history"] == ig["residual_history"] == ig2["residual_history"]
    assert np.array_equal(ue, ug) and np.array_equal(ue, ug2)
    assert any(isinstance(v, tuple) for v in sg._graphs.values())


@pytest.mark.parametrize("strategy", ["double", "adaptive"])
def test_full_multigrid_start_saves_cycles(strategy):
    n = 1025
    p = PoissonProblem.manufactured(n, on_device=True)
    u0, i0 = MixedPrecisionMultigrid(strategy).solve(p)
    u1, i1 = MixedPrecisionMultigrid(strategy, fmg=True).solve(p)
    assert i0["converged"] and i1["converged"] and i1["fmg"]
    assert i1["iterations"] <= i0["iterations"] - 2, (i0["iterations"], i1["iterations"])
    ref = O.mms_discretisation_error(n)
    for u in (u0, u1):
        assert abs(np.max(np.abs(u - O.mms_exact(n))) - ref) <= 0.01 * ref
    # the FMG iterate alone is already at discretisation-error level: first residual far below the zero-start one
    assert i1["residual_history"][0] < 0.05 * i0["residual_history"][0]


@pytest.mark.parametrize("strategy", ["adaptive", "double"])
def test_solve_many_pipelines_transfers_and_equals_sequential_solves(strategy):
    """solve_many: uploads / downloads of neighbouring solves overlap the cycles on side streams; every solution and
    residual history must equal the plain solve() of the same right-hand side bit for bit (5 solves > 2 staging slots)."""
    n = 257
    rng = np.random.default_rng(5)
    x = np.linspace(0, 1, n)
    X, Y = np.meshgrid(x, x, indexing="ij")
    rhs = [(k + 1) * np.sin((k + 1) * np.pi * X) * np.sin(np.pi * Y) + 0.1 * rng.standard_normal((n, n)) for k in range(5)]
    solver = MixedPrecisionMultigrid(precision_strategy=strategy, tolerance=1e-8)
    seq = []
    for f in rhs:
        u, info = solver.solve(PoissonProblem(rhs=f, nx=n, ny=n))
        seq.append((u.copy(), info["residual_history"]))
    pinned = []
    for f in rhs:
        t = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
        t.copy_(torch.from_numpy(f))
        pinned.append(PoissonProblem(rhs=t, nx=n, ny=n))
    for probs in (pinned, [PoissonProblem(rhs=f, nx=n, ny=n) for f in rhs]):  # pinned tensors, then plain NumPy
        sols, infos = solver.solve_many(probs)
        assert len(sols) == len(infos) == 5
        for (u, hist), u2, info in zip(seq, sols, infos):
            assert info["converged"] and info["residual_history"] == hist
            assert np.array_equal(u, u2)
    outs = [torch.empty((n, n), dtype=torch.float64, pin_memory=True) for _ in range(2)]
    sols, _ = solver.solve_many(pinned[:2], outputs=outs)
    assert np.array_equal(sols[1], seq[1][0]) and sols[1].ctypes.data == outs[1].numpy().ctypes.data
    with pytest.raises(ValueError, match="outputs"):
        solver.solve_many(pinned[:2], outputs=outs[:1])
    with pytest.raises(ValueError, match="share the grid"):
        solver.solve_many([pinned[0], PoissonProblem(rhs=np.zeros((129, 129)), nx=129, ny=129)])
    assert solver.solve_many([]) == ([], [])


def test_16385_mixed_headline_config_stops_on_the_rounding_floor_within_one_percent():
    """BASELINE configs[2].  The reference's absolute tolerance 1e-8 lies below the fp64 rounding floor of f - A u at
    h = 1/16384, so (solvers/policy.py) the solve stays in the fp32-cycle / fp64-residual refinement instead of switching
    to fp64 cycles at 1e-6, ends once the residual has not contracted for two consecutive cycles below the a-priori
    floor bound, and its MMS error must be within 1 % of the exactly converged discrete solution's (closed form,
    SURVEY 8c: 3.063928466e-9; measured history: profiles/r02_floor_study_16385.json).  ~10 GB of HBM, ~40 ms."""
    from mixed_precision_multigrid_solvers_for_pdes_b200 import ops
    n = 16385
    s = MixedPrecisionMultigrid("adaptive", switch_threshold=1e-6, tolerance=1e-8)
    s.setup(n, n)
    b64 = s._engine.levels[0].bufs(torch.float64)
    ops.fill_sinsin_(b64.f, (0.0, 1.0, 0.0, 1.0), 2 * np.pi ** 2, 1.0, 1.0)
    ops.zero_ring_(b64.f)
    u_dev, info = s._solve_device(False)       # the solve loop of solve(), without the 2 GB download
    hist = info["residual_history"]
    assert info["stopped_on"] == "rounding_floor" and not info["converged"]
    assert info["iterations"] == 10, hist
    assert hist[-1] > 0.5 * hist[-2] > 0.25 * hist[-3] and hist[-3] < 0.2 * hist[-4]   # contraction until the floor, then none
    assert 1e-8 < hist[-1] <= info["attainable_residual"] < 5e-7
    # the switch to fp64 cycles was due at cycle 7 (||r|| <= 1e-6) and skipped: the tolerance is below the floor bound
    assert info["precision_switches"] == [] and set(info["precision_history"]) == {"mixed"}
    assert info["switch_blocked"]["iteration"] == 7 and info["switch_blocked"]["residual"] <= 1e-6
    err = ops.maxerr_sinsin(u_dev)
    want = 3.063928466e-9
    assert abs(O.mms_discretisation_error(n) - want) < 1e-15
    assert abs(err - want) <= 0.01 * want, (err, want)
    del s, u_dev
    torch.cuda.empty_cache()


def test_floor_rule_never_fires_where_the_tolerance_is_attainable_and_ends_unattainable_solves():
    p = PoissonProblem(source_term, nx=257, ny=257)
    _, info = MixedPrecisionMultigrid("double", tolerance=1e-8).solve(p)
    assert info["stopped_on"] == "tolerance" and info["converged"] and info["iterations"] == 8
    _, low = MixedPrecisionMultigrid("double", tolerance=1e-15, max_iterations=40).solve(p)
    assert low["stopped_on"] == "rounding_floor" and low["iterations"] < 17
    assert low["final_residual"] <= low["attainable_residual"] < 1e-10
    _, off = MixedPrecisionMultigrid("double", tolerance=1e-15, max_iterations=14, stop_on_rounding_floor=False).solve(p)
    assert off["stopped_on"] is None and off["iterations"] == 14 and not off["converged"]


def test_solve_returns_an_array_the_caller_owns():
    n = 129
    s = MixedPrecisionMultigrid("double")
    x = np.linspace(0, 1, n)
    X, Y = np.meshgrid(x, x, indexing="ij")
    ua, _ = s.solve(PoissonProblem(rhs=np.sin(np.pi * X) * np.sin(np.pi * Y), nx=n, ny=n))
    keep = ua.copy()
    ub, _ = s.solve(PoissonProblem(rhs=3.0 * np.sin(2 * np.pi * X) * np.sin(np.pi * Y), nx=n, ny=n))
    assert np.array_equal(ua, keep) and not np.array_equal(ua, ub)     # the first result survived the second solve
    uv, _ = s.solve(PoissonProblem(rhs=np.sin(np.pi * X) * np.sin(np.pi * Y), nx=n, ny=n), reuse_output=True)
    assert np.array_equal(uv, keep) and uv.ctypes.data == s._pinned_out.numpy().ctypes.data  # opt-in zero-copy view


def test_two_solvers_do_not_share_reduction_scratch_under_graph_replay():
    """a per-device scratch that a larger solver outgrows must not be freed under the captured graphs of a
    smaller one.  Each engine owns its scratch; replays of the small solver stay bit-identical."""
    small = MixedPrecisionMultigrid("adaptive")
    ps = PoissonProblem(source_term, nx=513, ny=513)
    u0, i0 = small.solve(ps)
    u1, i1 = small.solve(ps)                     # graphs captured
    big = MixedPrecisionMultigrid("adaptive")
    big.solve(PoissonProblem(source_term, nx=2049, ny=2049))
    junk = [torch.full((1 << 20,), float("nan"), dtype=torch.float64, device="cuda") for _ in range(8)]
    u2, i2 = small.solve(ps)
    assert i0["residual_history"] == i1["residual_history"] == i2["residual_history"]
    assert np.array_equal(u0, u2)
    assert small._engine.workspace.data_ptr() != big._engine.workspace.data_ptr()
    del junk


def test_level_timings_are_filled_by_profile_levels():
    s = MixedPrecisionMultigrid("adaptive")
    p = PoissonProblem.manufactured(1025, on_device=True)
    s.solve(p)
    lt = s.profile_levels(cycles=2)
    _, info = s.solve(p)
    assert info["level_timings"] == lt and set(lt) == set(range(info["num_levels"]))
    assert lt[0]["smooth_time"] > 0 and lt[0]["passes"] >= 2 and lt[1]["smooth_time"] > 0
    assert lt[0]["smooth_time"] > lt[2]["smooth_time"]
