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
