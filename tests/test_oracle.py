"""The CPU oracle against the golden vectors produced by the unmodified reference (oracle/make_golden.py)."""
import numpy as np
import pytest

import gp_oracle as orc
from conftest import REAL_CONFIGS, load_golden

COND_OK = 1e8


def rel(a, b, scale=None):
    a, b = np.asarray(a), np.asarray(b)
    s = np.abs(b).max() if scale is None else scale
    return np.abs(a - b).max() / max(s, 1e-300)


def test_fixed_theta_numpy_restatement():
    g = load_golden("fixed_theta_synth")
    for m in g["sizes"]:
        t, y = g[f"t_{m}"], g[f"y_{m}"]
        for gi in range(2):
            for k, th in enumerate(g["thetas"]):
                if g[f"cond_{m}"][gi, k] > COND_OK:
                    continue
                lml, grad, st = orc.np_lml_grad(t, y[gi], th)
                assert st == 0
                assert abs(lml - g[f"lml_{m}"][gi, k]) <= 1e-10 * abs(g[f"lml_{m}"][gi, k])
                ref = g[f"grad_{m}"][gi, k]
                assert rel(grad, ref) <= 1e-9


def test_synthetic_generator_is_the_goldens_input():
    g = load_golden("fixed_theta_synth")
    t, y = orc.synthetic_trajectories(2, 200, seed=200)
    assert np.array_equal(t, g["t_200"]) and np.array_equal(y, g["y_200"])


@pytest.mark.parametrize("name", REAL_CONFIGS)
def test_real_config_fixed_theta(name):
    g = load_golden(name)
    T, Y = g["T"], g["Y"]
    for gi in range(T.shape[0]):
        for k, th in enumerate(g["thetas_eval"][gi]):
            if g["cond_eval"][gi, k] > COND_OK or not np.isfinite(g["lml_eval"][gi, k]):
                continue
            lml, grad, st = orc.np_lml_grad(T[gi], Y[gi], th)
            assert abs(lml - g["lml_eval"][gi, k]) <= 1e-10 * max(1.0, abs(g["lml_eval"][gi, k]))
            assert rel(grad, g["grad_eval"][gi, k], max(1.0, np.abs(g["grad_eval"][gi, k]).max())) <= 1e-8


@pytest.mark.parametrize("name", REAL_CONFIGS)
def test_real_config_moments(name):
    g = load_golden(name)
    T, Y, t_est, eta = g["T"], g["Y"], g["t_est"], float(g["eta"])
    for gi in range(min(T.shape[0], 3)):
        th = g["theta_opt"][gi]
        alpha, _ = orc.np_alpha(T[gi], Y[gi], th)
        assert rel(alpha, g["alpha_opt"][gi]) <= 1e-9
        mean, std = orc.np_predict(T[gi], Y[gi], th, t_est)
        assert rel(mean, g["pred_mean"][gi]) <= 1e-10
        assert rel(std, g["pred_std"][gi]) <= 1e-8
        out = orc.np_lstsq_moments(T[gi], Y[gi], th, t_est, eta, want_sqrtW=False)
        assert rel(out["state_estimate"], g["state_estimate"][gi]) <= 1e-10
        assert rel(out["ddt_estimate"], g["ddt_estimate"][gi]) <= 1e-10
        if gi < g["ddt_covariance"].shape[0]:
            assert rel(out["ddt_covariance"], g["ddt_covariance"][gi]) <= 1e-9


def test_oracle_gp_port_reproduces_reference_fit():
    """OracleGP (sklearn-driven port of GP_RBFW) with the reference's RNG contract reproduces the golden optimum."""
    g = load_golden("heat_1_20_05_80_5")
    b = g["bounds"]
    gi = 0
    # replay the restart points through the global RNG: OracleGP draws them exactly like the reference
    class _Replay:
        def __init__(self, pts):
            self.pts, self.k = pts, 0

        def uniform(self, lo, hi):
            p = self.pts[self.k]
            self.k += 1
            return p

    gp = orc.OracleGP(tuple(b[0]), tuple(b[1]), tuple(b[2]), int(g["n_restarts"]))
    gp.gpr.random_state = None
    import sklearn.utils

    orig = sklearn.gaussian_process._gpr.check_random_state
    sklearn.gaussian_process._gpr.check_random_state = lambda rs: _Replay(g["starts"][gi])
    try:
        gp.fit(g["T"][gi], g["Y"][gi])
    finally:
        sklearn.gaussian_process._gpr.check_random_state = orig
    assert abs(gp.lml - g["lml_opt"][gi]) <= 1e-9 * abs(g["lml_opt"][gi])
    assert np.allclose(gp.theta, g["theta_opt"][gi], atol=1e-4)


def test_reference_import_matches_oracle():
    import ref_import

    if not ref_import.reference_available():
        pytest.skip("/root/reference not present (GPU box)")
    gpk = ref_import.load_reference_gpkernels()
    t, y = orc.synthetic_trajectories(1, 60, seed=3)
    b = np.array([(1e-5, 1e5), (1e-5, 1e2), (1e-16, 1e2)])
    gp = gpk.GP_RBFW(tuple(b[0]), tuple(b[1]), tuple(b[2]), 0)
    gp.gpr.optimizer = None
    gp.fit(t, y[0])
    th = np.log([1.3, 0.08, 2e-3])
    l_ref, g_ref = gp.gpr.log_marginal_likelihood(th, eval_gradient=True)
    l, gr, _ = orc.np_lml_grad(t, y[0], th)
    assert abs(l - l_ref) <= 1e-11 * abs(l_ref)
    assert rel(gr, g_ref) <= 1e-9
    gp.gpr.kernel = gp.gpr.kernel.clone_with_theta(th)
    gp.gpr.fit(t[:, None], y[0])  # optimizer None: fit() only refreshes kernel_, L_, alpha_ at theta
    t_est = np.linspace(0, 1, 50)
    gp.compute_lstsq_matrices(t_est, eta=1e-8)
    out = orc.np_lstsq_moments(t, y[0], th, t_est, 1e-8, want_sqrtW=False)
    assert rel(out["ddt_covariance"], gp.ddt_covariance) <= 1e-10
    assert rel(out["ddt_estimate"], gp.ddt_estimate) <= 1e-11


# ------------------------------------------------------------------ Matern extension of the oracle
@pytest.mark.parametrize("twice_nu", [3, 5])
def test_matern_oracle_formulas(twice_nu):
    """np_matern against scikit-learn's Matern, its derivatives against central differences."""
    rng = np.random.default_rng(twice_nu)
    t1, t2 = np.sort(rng.uniform(0, 1, 40)), np.sort(rng.uniform(0, 1, 31))
    th = np.log([1.3, 0.2, 1e-3])
    sk = orc.sk_matern_kernel(th, twice_nu)
    assert rel(orc.np_matern(t1, t2, 1.3, 0.2, twice_nu), sk(t1[:, None], t2[:, None])) <= 1e-14
    h = 1e-5
    k = lambda a, b: orc.np_matern(a, b, 1.3, 0.2, twice_nu)
    d1 = (k(t1 + h, t2) - k(t1 - h, t2)) / (2 * h)
    far = np.abs(t1[:, None] - t2[None, :]) > 10 * h
    assert np.abs(d1 - orc.np_matern(t1, t2, 1.3, 0.2, twice_nu, deriv=1))[far].max() <= 1e-7 * np.abs(d1).max()
    d2 = (k(t1 + h, t2 + h) - k(t1 + h, t2 - h) - k(t1 - h, t2 + h) + k(t1 - h, t2 - h)) / (4 * h * h)
    assert np.abs(d2 - orc.np_matern(t1, t2, 1.3, 0.2, twice_nu, deriv=2))[far].max() <= 1e-4 * np.abs(d2).max()
    # LML / gradient restatement against sklearn's own regressor
    y = np.sin(7 * t1) + 0.05 * rng.standard_normal(40)
    gp = orc.OracleGP((1e-5, 1e5), (1e-5, 1e2), (1e-16, 1e2), 0, twice_nu=twice_nu)
    gp.gpr.optimizer = None
    gp.fit(t1, y)
    l_ref, g_ref = gp.lml_grad(th)
    l, g, st = orc.np_lml_grad_matern(t1, y, th, twice_nu)
    assert st == 0 and abs(l - l_ref) <= 1e-12 * abs(l_ref) and rel(g, g_ref) <= 1e-10
