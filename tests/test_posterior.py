"""Step-3 posterior assembly for a grid of regularizers: the CUDA path (gpbo_posterior_grid_host)
against the oracle restatement of PDEs/step3_estimate.py:75-95 + codebase/wlstsq.py:183-188 (oracle/gp_oracle.py::
np_posterior_grid; PARITY UNPINNED -- the reference's step 3 needs `opinf`, which is not installed), on the reference's
own Euler / SEIRD moments (golden fixtures) and on synthetic weights.  Tolerances: 1e-10 relative for the precision data
(gram, Cholesky factor), 1e-9 relative for the means (conditioning of the regularised least squares)."""
import numpy as np
import pytest

from conftest import load_golden, record

import gp_oracle as orc


def quadratic_data_matrix(Q):
    """[1 | q^T | compact q (x) q] -- the shape of opinf's data matrix for a 'cAH' model (any D serves the test)."""
    r, n = Q.shape
    iu = np.triu_indices(r)
    return np.hstack([np.ones((n, 1)), Q.T, (Q.T[:, :, None] * Q.T[:, None, :])[:, iu[0], iu[1]]])


def reference_sqrtW(C, eta):
    """gpkernels.py:496-504 (the eigh route of the reference)."""
    out = []
    for Ci in C:
        lam, V = np.linalg.eigh(Ci + eta * np.eye(Ci.shape[0]))
        assert lam.min() > 0
        out.append(V @ np.diag(lam ** -0.5) @ V.T)
    return np.array(out)


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


# ---------------------------------------------------------------- oracle (CPU)
def test_oracle_posterior_grid_satisfies_its_defining_equations():
    rng = np.random.default_rng(0)
    r, n, d = 3, 60, 9
    M = rng.standard_normal((r, n, n))
    W = np.array([m @ m.T / n + 0.1 * np.eye(n) for m in M])
    D, Z = rng.standard_normal((n, d)), rng.standard_normal((r, n))
    regs = np.array([0.0, 1e-2, 3.0])
    o = orc.np_posterior_grid(W, D, Z, regs)
    for k, reg in enumerate(regs):
        for i in range(r):
            A, b = W[i] @ D, W[i] @ Z[i]
            P = A.T @ A + reg**2 * np.eye(d)
            assert rel(P @ o["means"][k, i], A.T @ b) <= 1e-11                       # normal equations
            assert rel(o["chol"][k, i] @ o["chol"][k, i].T, P) <= 1e-13 and o["status"][k, i] == 0
            aug = np.vstack([A, reg * np.eye(d)])                                    # the stacked form of wlstsq.py:118
            x = np.linalg.lstsq(aug, np.concatenate([b, np.zeros(d)]), rcond=None)[0]
            assert rel(o["means"][k, i], x) <= 1e-10
    # a zero column with reg = 0: singular precision -> the reference's "Matrix is not positive definite"
    D0 = D.copy()
    D0[:, 4] = 0.0
    o0 = orc.np_posterior_grid(W, D0, Z, np.array([0.0, 1.0]))
    assert np.all(o0["status"][0] == 1) and np.all(o0["status"][1] == 0)


def test_posterior_grid_host_mirror_logic():
    """PosteriorGrid (the host object the reference-side loop reads) on the oracle's numbers: precisions as in
    step3_estimate.py:84-90, the not-positive-definite flag, draws distributed as N(mean, precision^-1)."""
    from gpbo_pkg import pkg

    rng = np.random.default_rng(4)
    r, n, d = 2, 40, 5
    M = rng.standard_normal((r, n, n))
    W = np.array([m @ m.T / n + 0.2 * np.eye(n) for m in M])
    D, Z = rng.standard_normal((n, d)), rng.standard_normal((r, n))
    regs = np.array([0.0, 2.0])
    pg = pkg.step3_posterior.PosteriorGrid(regs, orc.np_posterior_grid(W, D, Z, regs))
    A0 = W[0] @ D
    assert rel(pg.precisions(1)[0], A0.T @ A0 + 4.0 * np.eye(d)) <= 1e-13 and pg.is_spd(0) and pg.is_spd(1)
    draws = np.array([pg.draw(1, rng)[0] for _ in range(4000)])
    cov = np.linalg.inv(pg.precisions(1)[0])
    assert np.abs(draws.mean(0) - pg.means[1, 0]).max() <= 5 * np.sqrt(cov.diagonal().max() / 4000)
    assert rel(np.cov(draws.T), cov) <= 0.15
    pg.status[0, 1] = 1
    assert not pg.is_spd(0)
    with pytest.raises(np.linalg.LinAlgError):
        pg.draw(0, rng)


# ---------------------------------------------------------------- CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["euler_006_200_03_400_6", "seird_090_090_10_360"])
def test_posterior_grid_real_moments(ctx, name):
    if ctx.path != "blocked":
        pytest.skip("path independent")
    g = load_golden(name)
    Q, Z, C, eta = g["state_estimate"], g["ddt_estimate"], g["ddt_covariance"], float(g["eta"])
    rc = C.shape[0]
    W = reference_sqrtW(C, eta)
    D = quadratic_data_matrix(Q)
    # The quadratic data matrix of the SEIRD states is rank deficient (S + E + I + R + D = 1: cond(D) = 3e15), so its
    # grid starts where the regularised problem is defined to double precision (cond(P) <= 2e13); Euler: cond(A) = 1.6e3.
    well = name.startswith("euler")
    regs = np.logspace(-4, 4, 17) if well else np.logspace(-2, 4, 13)
    ref = orc.np_posterior_grid(W, D, Z[:rc], regs)
    got = ctx.posterior_grid(D, Z[:rc], regs, sqrtW=W)
    assert np.array_equal(got["status"], ref["status"]) and not got["status"].any()
    e_gram, e_proj = rel(got["gram"], ref["gram"]), rel(got["proj"], ref["proj"])
    e_mean = max(rel(got["means"][k, i], ref["means"][k, i]) for k in range(regs.size) for i in range(rc))
    # Cholesky factors: backward error (C C^T against the precision of step3_estimate.py:86-90) everywhere; element-wise
    # against LAPACK's factor where the factor itself is well conditioned
    eye = np.eye(D.shape[1])
    e_chol_back = max(rel(got["chol"][k, i] @ got["chol"][k, i].T, got["gram"][i] + regs[k] ** 2 * eye)
                      for k in range(regs.size) for i in range(rc))
    e_chol = max(rel(got["chol"][k], ref["chol"][k]) for k in range(regs.size))
    for key, v in (("gram", e_gram), ("proj", e_proj), ("chol_backward", e_chol_back), ("mean", e_mean)):
        record(f"posterior_grid_{key}_rel[{name}]", v)
    assert e_gram <= 1e-10 and e_proj <= 1e-10 and e_chol_back <= 1e-13 and e_mean <= 1e-9
    if well:
        record(f"posterior_grid_chol_rel[{name}]", e_chol)
        assert e_chol <= 1e-10
    for k in (0, regs.size - 1):
        for i in range(rc):
            assert np.all(np.triu(got["chol"][k, i], 1) == 0.0)      # strict upper part of the factor is zero


@pytest.mark.gpu
def test_posterior_grid_resident_weights_and_host_mirror(ctx):
    """The weight matrices stay in HBM after compute_lstsq_matrices: sqrtW=None uses them; the host mirror
    (step3_posterior.posterior_grid) returns the precisions list and draws like scipy's from_precision object."""
    if ctx.path != "blocked":
        pytest.skip("path independent")
    from gpbo_pkg import pkg

    g = load_golden("euler_006_050_01_400_6")
    T, Y, theta, t_est, eta = g["T"], g["Y"], g["theta_opt"], g["t_est"], float(g["eta"])
    state, ddt, cov, w, st, wst, _ = ctx.lstsq_weights(T, Y, theta, t_est, eta)
    assert not st.any() and not wst.any()
    D = quadratic_data_matrix(state)
    regs = np.array([1e-2, 1.0, 50.0])
    a = ctx.posterior_grid(D, ddt, regs)                        # resident weights
    b = ctx.posterior_grid(D, ddt, regs, sqrtW=w)               # the same matrices through the host
    for key in ("means", "chol", "gram", "proj"):
        assert np.array_equal(a[key], b[key]), key
    ref = orc.np_posterior_grid(w, D, ddt, regs)
    assert max(rel(a["means"][k], ref["means"][k]) for k in range(3)) <= 1e-9
    pg = pkg.step3_posterior.posterior_grid(None, D, ddt, regs, ctx=ctx)
    assert pg.is_spd(1) and len(pg.precisions(1)) == T.shape[0]
    assert rel(pg.precisions(1)[0], ref["gram"][0] + np.eye(D.shape[1])) <= 1e-10
    rng = np.random.default_rng(3)
    draws = np.array([pg.draw(1, rng) for _ in range(400)])
    cov0 = np.linalg.inv(pg.precisions(1)[0])
    assert np.abs(draws[:, 0].mean(0) - pg.means[1, 0]).max() <= 6 * np.sqrt(np.diag(cov0).max() / 400)


@pytest.mark.gpu
@pytest.mark.parametrize("n,d", [(3200, 128), (50, 1), (97, 17)])
def test_posterior_grid_sizes(ctx, n, d):
    """Largest supported operator row (d = 128: 135 KB of dynamic shared memory) at the reference's m' = 3200, a single
    unknown, and sizes off every tile."""
    if ctx.path != "blocked":
        pytest.skip("path independent")
    rng = np.random.default_rng(n + d)
    r = 2
    # weights with the structure of (C + eta I)^(-1/2): smooth SPD matrices, built without an n^3 eigen-decomposition
    t = np.linspace(0, 1, n)
    W = np.array([np.exp(-0.5 * ((t[:, None] - t[None, :]) / ell) ** 2) + 0.5 * np.eye(n) for ell in (0.05, 0.2)])
    D, Z = rng.standard_normal((n, d)), rng.standard_normal((r, n))
    regs = np.array([1e-3, 1.0, 30.0])
    got, ref = ctx.posterior_grid(D, Z, regs, sqrtW=W), orc.np_posterior_grid(W, D, Z, regs)
    assert not got["status"].any() and not ref["status"].any()
    assert rel(got["gram"], ref["gram"]) <= 1e-11 and rel(got["proj"], ref["proj"]) <= 1e-11
    e_mean = max(rel(got["means"][k, i], ref["means"][k, i]) for k in range(3) for i in range(r))
    e_chol = max(rel(got["chol"][k, i], ref["chol"][k, i]) for k in range(3) for i in range(r))
    record(f"posterior_grid_mean_rel[n={n}, d={d}]", e_mean)
    assert e_mean <= 1e-9 and e_chol <= 1e-9


@pytest.mark.gpu
def test_posterior_grid_not_positive_definite_and_errors(ctx):
    if ctx.path != "blocked":
        pytest.skip("path independent")
    from gpbo_pkg import pkg

    rng = np.random.default_rng(1)
    r, n, d = 2, 130, 33
    M = rng.standard_normal((r, n, n))
    W = np.array([m @ m.T / n + 0.05 * np.eye(n) for m in M])
    D, Z = rng.standard_normal((n, d)), rng.standard_normal((r, n))
    D[:, 7] = 0.0                                                # singular Gram matrix
    regs = np.array([0.0, 0.5])
    got, ref = ctx.posterior_grid(D, Z, regs, sqrtW=W), orc.np_posterior_grid(W, D, Z, regs)
    assert np.array_equal(got["status"], ref["status"]) and np.all(got["status"][0] == 1)
    assert np.all(np.isnan(got["means"][0])) and rel(got["means"][1], ref["means"][1]) <= 1e-9
    with pytest.raises(ValueError):
        ctx.posterior_grid(D[:-1], Z, regs, sqrtW=W)
    with pytest.raises(ValueError):
        ctx.posterior_grid(D, Z, [np.nan], sqrtW=W)
    with pytest.raises(pkg.GpboError):
        ctx.posterior_grid(rng.standard_normal((n, 129)), Z, regs, sqrtW=W)      # d > 128
