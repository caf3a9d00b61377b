"""Host-side logic that needs no GPU: Grid geometry, PrecisionManager rules, argument validation.
Expected values are the reference's documented behaviour (SURVEY.md sections 4 and 8a)."""
import numpy as np
import pytest

from mixed_precision_multigrid_solvers_for_pdes_b200 import (Grid, LaplacianOperator, PrecisionLevel, PrecisionManager,
                                                             ProlongationOperator, RestrictionOperator)


def test_grid_geometry():
    g = Grid(129, 65, (0.0, 2.0, -1.0, 1.0))
    assert g.shape == (129, 65) and g.size == 129 * 65
    assert g.hx == 2.0 / 128 and g.hy == 2.0 / 64 and g.h == min(g.hx, g.hy)
    c = g.coarsen()
    assert c.shape == (65, 33) and c.domain == g.domain
    assert g.refine().shape == (257, 129)
    assert g.X.shape == (129, 65) and g.X[3, 0] == g.x[3] and g.Y[0, 5] == g.y[5]
    with pytest.raises(ValueError):
        Grid(2, 5)
    with pytest.raises(ValueError):
        Grid(10, 9).coarsen()


def test_grid_l2_norm_numpy_matches_reference_formula():
    g = Grid(9, 17)
    f = np.random.default_rng(0).standard_normal((9, 17))
    assert g.l2_norm(f) == np.sqrt(g.hx * g.hy * np.sum(f ** 2))
    g.apply_dirichlet_bc(3.0)
    assert np.all(g.values[0, :] == 3.0) and np.all(g.values[:, -1] == 3.0) and g.values[4, 4] == 0.0


def test_precision_thresholds():
    # reference tests/unit/test_precision.py:105-117, 148-167
    pm = PrecisionManager("double", adaptive=True, convergence_threshold=1e-6)
    assert pm.should_downgrade_precision([(9, 9)], 1e-3) and not pm.should_downgrade_precision([(9, 9)], 1e-5)
    assert pm.update_precision(1e-3, [(9, 9)]) and pm.current_precision == PrecisionLevel.SINGLE
    assert not pm.update_precision(1e-4, [(9, 9)])
    assert pm.update_precision(5e-6, [(9, 9)]) and pm.current_precision == PrecisionLevel.DOUBLE
    assert [p.value for p in pm.precision_history] == ["float64", "float32", "float64"]
    mixed = PrecisionManager("mixed")
    assert [mixed.get_precision_for_level(l, 4) for l in range(4)] == [PrecisionLevel.DOUBLE] * 2 + [PrecisionLevel.SINGLE] * 2
    fixed = PrecisionManager("single", adaptive=False)
    assert not fixed.update_precision(1e-9, [(9, 9)]) and fixed.get_dtype() == np.float32
    with pytest.raises(ValueError):
        PrecisionManager("half")


def test_precision_memory_rule_and_promotion():
    pm = PrecisionManager("double", memory_threshold_gb=1e-6)
    assert pm.should_downgrade_precision([(1025, 1025)], 1e-12)
    pm2 = PrecisionManager("single")
    assert pm2.should_promote_precision([1.0, 0.99, 0.985, 0.98, 0.979], PrecisionLevel.SINGLE)
    assert not pm2.should_promote_precision([1.0, 0.1, 0.01, 1e-3, 1e-4], PrecisionLevel.SINGLE)
    assert not pm2.should_promote_precision([1.0, 0.99, 0.985, 0.98, 0.979], PrecisionLevel.DOUBLE)


def test_operator_validation_messages():
    # reference tests/unit/test_operators.py:156,237 match on these messages
    with pytest.raises(ValueError, match="Unknown restriction method"):
        RestrictionOperator("cubic")
    with pytest.raises(ValueError, match="Unknown prolongation method"):
        ProlongationOperator("cubic")
    g = Grid(9, 9)
    with pytest.raises(ValueError, match="doesn't match grid shape"):
        LaplacianOperator().apply(g, np.zeros((8, 9)))
    with pytest.raises(ValueError, match="Cannot restrict"):
        RestrictionOperator().apply(g, np.zeros((9, 9)), Grid(9, 9))
    with pytest.raises(ValueError, match="Cannot prolongate"):
        ProlongationOperator().apply(Grid(5, 5), np.zeros((5, 5)), Grid(11, 11))
    assert LaplacianOperator(2.0).apply_stencil(g, np.ones((9, 9)), 3, 3) == 0.0


# ---- solvers/policy.py: the stopping / switching rules every driver shares (pure Python) -----------------------------
def _policy(mode, **kw):
    import importlib.util
    import os
    # load the module by path: the package __init__ imports torch + the CUDA library, the policy needs neither
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "mixed_precision_multigrid_solvers_for_pdes_b200", "solvers", "policy.py")
    spec = importlib.util.spec_from_file_location("mgb200_policy", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    h = kw.pop("h", 1.0 / 128)
    args = dict(tolerance=1e-8, switch_threshold=1e-6, hx=h, hy=h)
    args.update(kw)
    return mod, mod.CyclePolicy(mode, **args)


def test_policy_reference_tolerance_and_switch_threshold():
    mod, p = _policy("switch", u_norm=lambda ph: 0.5)
    hist = [7.562e-1, 4.097e-2, 2.219e-3, 1.202e-4, 6.513e-6, 3.528e-7, 1.911e-8, 1.036e-9]  # SURVEY 8c, 129^2
    phases, actions = [], []
    for r in hist:
        phases.append(p.phase)
        actions.append(p.observe(r))
    assert actions == [mod.CONTINUE] * 7 + [mod.CONVERGED] and p.stopped_on == "tolerance"
    assert phases == ["refine"] * 6 + ["fp64"] * 2
    assert p.switches == [{"iteration": 6, "residual": 3.528e-7, "from": "mixed", "to": "float64",
                           "reason": "switch_threshold"}]
    assert p.floor_bound == 2.220446049250313e-16 * 4 * 128 ** 2 * 0.5  # evaluated once, at the switch decision
    assert p.switch_blocked is None                                   # tolerance 1e-8 >> bound 7e-12: switch as documented


def test_policy_rounding_floor_needs_stagnation_and_the_a_priori_bound():
    # 16385^2: the fp64 floor of f - A u is ~3e-8 > 1e-8; history of the round-1 run + the cycle that detects the floor
    h = 1.0 / 16384
    calls = []
    mod, p = _policy("switch", h=h, u_norm=lambda ph: calls.append(1) or 0.5)
    hist = [17.87, 1.148, 6.35e-2, 3.44e-3, 1.88e-4, 9.75e-6, 5.68e-7, 2.87e-8, 2.81e-8, 2.83e-8]
    actions = [p.observe(r) for r in hist]
    assert actions == [mod.CONTINUE] * 9 + [mod.FLOOR] and p.stopped_on == "rounding_floor"
    assert len(calls) == 1                                   # one iterate norm per solve, when the bound is first needed
    # the tolerance (1e-8) lies below the floor bound (1.2e-7): the solve stays in the refinement, whose smooth-error
    # contraction survives on the floor (profiles/r02_floor_study_16385.json), instead of switching to fp64 at 1e-6
    assert p.switches == [] and p.phase == "refine" and p.switch_blocked["iteration"] == 7
    # ... where the tolerance is attainable the switch happens as documented (4097^2: bound 7.4e-9 < 1e-8)
    mod, sw = _policy("switch", h=1.0 / 4096, u_norm=lambda ph: 0.5)
    assert [sw.observe(r) for r in (7.57e-1, 4.1e-2, 2.2e-3, 1.2e-4, 6.5e-6, 3.6e-7, 1.9e-8, 1.1e-9)][-1] == mod.CONVERGED
    assert sw.switches[0]["iteration"] == 6 and sw.switch_blocked is None
    mod, one = _policy("switch", h=h, u_norm=lambda ph: 0.5, floor_confirmations=1)
    assert [one.observe(r) for r in hist[:9]][-1] == mod.FLOOR
    assert abs(one.floor_bound - 2.220446049250313e-16 * 4 * 16384 ** 2 * 0.5) < 1e-20
    assert hist[-1] <= p.floor_bound
    # a slowly converging solve far ABOVE the bound is not mistaken for the floor
    mod, q = _policy("fp64", h=1.0 / 128, u_norm=lambda ph: 0.5)
    assert [q.observe(r) for r in (1.0, 0.7, 0.5, 0.36)] == [mod.CONTINUE] * 4 and q.stopped_on is None
    # switched off: runs on
    mod, off = _policy("switch", h=h, u_norm=lambda ph: 0.5, stop_on_floor=False)
    assert [off.observe(r) for r in hist][-1] == mod.CONTINUE
    # the fp32-only strategy runs on like the reference's all-fp32 solves
    mod, single = _policy("fp32", h=1.0 / 128, u_norm=lambda ph: 0.5)
    assert [single.observe(r) for r in (1.0, 1e-2, 9.6e-4, 9.5e-4, 9.5e-4, 9.5e-4)] == [mod.CONTINUE] * 6


def test_policy_stagnating_refinement_is_promoted_then_ends_on_the_floor():
    mod, p = _policy("refine", h=1.0 / 16384, u_norm=lambda ph: 0.5)
    # refinement mode never switches on the threshold; two ratios > 0.95 above the floor bound promote to fp64
    for r in (1.0, 1e-2, 9.9e-3, 9.8e-3):
        assert p.observe(r) == mod.CONTINUE
    assert p.phase == "fp64" and p.switches[0]["reason"] == "stagnation"
    assert p.observe(3e-8) == mod.CONTINUE and p.observe(2.9e-8) == mod.CONTINUE and p.observe(2.9e-8) == mod.FLOOR
    # ... and a refinement whose fp64 residual already sits on the floor stops without the detour
    mod, q = _policy("refine", h=1.0 / 16384, u_norm=lambda ph: 0.5)
    acts = [q.observe(r) for r in (1.0, 1e-3, 3.0e-8, 2.95e-8, 2.9e-8)]
    assert acts[-1] == mod.FLOOR and q.phase == "refine"
    with __import__("pytest").raises(ValueError):
        _policy("half")


def test_policy_last_cycle_hint():
    """CyclePolicy.likely_last: true before a cycle that should meet the tolerance at the last contraction rate, and
    before the confirming cycle on the rounding floor; never in the fp64 phase."""
    from mixed_precision_multigrid_solvers_for_pdes_b200.solvers.policy import CONTINUE, CyclePolicy
    p = CyclePolicy("refine", 1e-8, 1e-6, 1 / 1024, 1 / 1024, u_norm=lambda phase: 0.5)
    assert not p.likely_last()
    for norm, expect in ((1e-1, False), (1e-2, False), (1e-3, False), (1e-4, False), (1e-5, False), (1e-6, False),
                         (5e-8, True)):
        assert p.observe(norm) == CONTINUE
        assert p.likely_last() == expect, norm
    # rounding floor: h = 1/16384 -> bound 1.2e-7 * ||u||; the first stagnating cycle arms the hint
    q = CyclePolicy("refine", 1e-8, 1e-6, 1 / 16384, 1 / 16384, u_norm=lambda phase: 0.5)
    for norm in (1e-2, 1e-3, 1e-4, 1e-5, 1e-6, 1.2e-7):
        assert q.observe(norm) == CONTINUE
    assert not q.likely_last()
    assert q.observe(6.5e-8) == CONTINUE and q._floor_hits == 1 and q.likely_last()
    assert q.observe(5.9e-8) == "rounding_floor"
    r = CyclePolicy("fp64", 1e-8, 1e-6, 1 / 1024, 1 / 1024)
    r.observe(1e-3), r.observe(5e-8)
    assert not r.likely_last()
