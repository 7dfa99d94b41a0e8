"""Host-only tests of the optimiser pool C-ABI (gpbo_optpool_*): no CUDA device needed."""
import numpy as np
import pytest

from gpbo_pkg import pkg

BL = np.array([[-3.0, 3.0], [-2.0, 2.0], [-5.0, 1.0]])


def _objective(theta, c):
    d = theta - c
    return -(d ** 2).sum(1) - 0.1 * (d ** 4).sum(1), -2 * d - 0.4 * d ** 3


def test_pool_matches_single_minimiser():
    """The pool advances every pair exactly like gpbo_lbfgsb_minimize runs one (same state machine)."""
    rng = np.random.default_rng(0)
    starts = rng.uniform(BL[:, 0], BL[:, 1], size=(7, 3))
    c = rng.uniform(-1, 1, size=(7, 3))
    pool = pkg._lib.OptimizerPool(BL, starts)
    rounds = 0
    while True:
        idx, th = pool.live()
        if idx.size == 0:
            break
        lml, grad = _objective(th, c[idx])
        pool.feed(idx, lml, grad)
        rounds += 1
    res = pool.result()
    assert res["rounds"] == rounds and res["evals"] == int(res["nfev"].sum())
    for k in range(7):
        one = pkg._lib.lbfgsb_minimize(lambda x, k=k: tuple(-v for v in map(lambda a: a[0], _objective(x[None], c[k:k + 1]))),
                                       starts[k], BL)
        assert np.array_equal(one["x"], res["theta"][k]) and one["fun"] == res["fun"][k]
        assert one["nfev"] == res["nfev"][k] and one["status"] == res["status"][k]


def test_pool_early_stop_reports_last_accepted_iterate():
    starts = np.array([[2.5, 1.5, -4.0], [0.1, 0.1, 0.1]])
    c = np.zeros((2, 3))
    pool = pkg._lib.OptimizerPool(BL, starts)
    for _ in range(2):
        idx, th = pool.live()
        lml, grad = _objective(th, c[idx])
        pool.feed(idx, lml, grad)
    res = pool.result()
    assert np.all(res["status"] == -1)
    f0 = -_objective(starts, c)[0]
    assert np.all(res["fun"] <= f0) and np.all(np.isfinite(res["theta"]))
    # a pair that was never evaluated has no value yet
    fresh = pkg._lib.OptimizerPool(BL, starts).result()
    assert np.all(fresh["status"] == -1) and np.all(np.isinf(fresh["fun"])) and np.array_equal(fresh["theta"], starts)


def test_pool_rejects_bad_arguments():
    with pytest.raises(ValueError):
        pkg._lib.OptimizerPool(BL, np.zeros((3, 2)))
    with pytest.raises(pkg.GpboError):
        pkg._lib.OptimizerPool(np.array([[1.0, 0.0], [0, 1], [0, 1]]), np.zeros((1, 3)))     # empty interval
    pool = pkg._lib.OptimizerPool(BL, np.zeros((2, 3)))
    with pytest.raises(ValueError):
        pool.feed(np.array([0]), np.zeros(2), np.zeros((2, 3)))
    idx, th = pool.live()
    pool.feed(idx, np.array([-np.inf, -np.inf]), np.zeros((2, 3)))          # objective not finite at the start
    assert np.all(pool.result()["status"] == 5)
    with pytest.raises(pkg.GpboError):
        pool.feed(np.array([0], dtype=np.int32), np.zeros(1), np.zeros((1, 3)))           # pair no longer running


def test_shape_validation_before_the_library_is_called():
    """inconsistent (t, y, theta, gp_of) must raise ValueError instead of reading out of bounds."""
    c = object.__new__(pkg._lib.Context)          # no device: validation happens before any library call
    with pytest.raises(ValueError, match="inconsistent numbers of samples"):
        pkg._lib.Context.fit(c, np.zeros((1, 5)), np.zeros((1, 4)), BL, np.zeros((2, 3)), np.zeros(2, np.int32))
    with pytest.raises(ValueError, match="theta must have shape"):
        pkg._lib.Context.lml_grad(c, np.zeros((1, 5)), np.zeros((1, 5)), np.zeros((2, 2)))
    with pytest.raises(ValueError, match="gp_of"):
        pkg._lib.Context.lml_grad(c, np.zeros((2, 5)), np.zeros((2, 5)), np.zeros((3, 3)), np.array([0, 1]))
    with pytest.raises(ValueError, match="out of range"):
        pkg._lib.Context.lml_grad(c, np.zeros((2, 5)), np.zeros((2, 5)), np.zeros((3, 3)), np.array([0, 1, 2]))
    with pytest.raises(ValueError, match="theta must have shape"):
        pkg._lib.Context.predict(c, np.zeros((2, 5)), np.zeros((2, 5)), np.zeros((3, 3)), np.zeros(4))
