"""CPU oracle for the step2_fitgps hot path  --  TEST INFRASTRUCTURE ONLY.

This module is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  Nothing under ``gp-bayesopinf_b200/`` imports
it, and the product path raises when its CUDA library is missing.

Two independent restatements of the reference algorithm live here:

* ``np_*`` functions: plain NumPy/SciPy-LAPACK formulas, each citing the
  reference (``/root/reference/codebase/gpkernels.py``) or the third-party code
  the reference delegates to (scikit-learn ``gaussian_process/_gpr.py`` /
  ``kernels.py``; pinned by the reference at scikit-learn==1.5.2,
  scipy==1.14.1, numpy==2.1.3 -- ``requirements.txt:4-7``; this image carries
  scikit-learn 1.9.0 / scipy 1.18.1 / numpy 2.3.5, recorded in every golden
  file).
* ``OracleGP``: a restatement of ``gpkernels.GP_RBFW`` (``gpkernels.py:299-649``)
  that, like the reference, drives scikit-learn's ``GaussianProcessRegressor``
  (alpha=0, kernel ``C*RBF + White``) and SciPy's L-BFGS-B.  It is what
  ``bench.py`` times as the CPU baseline (kind "port").

Parity pinning: the reference ships no tests or golden vectors,
so the oracle is pinned against outputs of the *unmodified* reference
``GP_RBFW`` executed in the build container (``oracle/ref_import.py`` +
``oracle/make_golden.py`` -> ``tests/golden/*.npz``); ``tests/test_oracle.py``
checks both restatements against those fixtures.
"""

from __future__ import annotations

import numpy as np
import scipy.linalg as la

LOG_2PI = float(np.log(2.0 * np.pi))


# --------------------------------------------------------------------------
# Plain NumPy restatement of the arithmetic
# --------------------------------------------------------------------------
def np_kernel(t, theta, t2=None, eval_gradient=False):
    """K(theta) = sigma^2 * R + chi * I and (optionally) dK/dtheta.

    Follows sklearn ``kernels.py``: RBF ``:1559-1578`` (X/ell, squared
    euclidean distance, exp(-d/2), diagonal forced to 1), ConstantKernel
    ``:1279-1292``, Product ``:966-969``, WhiteKernel ``:1407-1419`` (zero when
    a second argument is given), Sum ``:866-869``.  ``theta`` is the log-space
    vector (log sigma^2, log ell, log chi) -- ``kernels.py:290-362``.
    """
    sig2, ell, chi = np.exp(np.asarray(theta, dtype=np.float64))
    x = np.asarray(t, dtype=np.float64) / ell
    if t2 is None:
        d = (x[:, None] - x[None, :]) ** 2
        R = np.exp(-0.5 * d)
        np.fill_diagonal(R, 1.0)
        K = sig2 * R
        K[np.diag_indices_from(K)] += chi
        if not eval_gradient:
            return K
        dK = np.empty(K.shape + (3,))
        dK[..., 0] = sig2 * R
        dK[..., 1] = sig2 * (R * d)
        dK[..., 2] = chi * np.eye(K.shape[0])
        return K, dK
    x2 = np.asarray(t2, dtype=np.float64) / ell
    d = (x[:, None] - x2[None, :]) ** 2
    return sig2 * np.exp(-0.5 * d)


def np_lml_grad(t, y, theta):
    """Log marginal likelihood and its gradient w.r.t. log-hyperparameters.

    Follows ``GaussianProcessRegressor.log_marginal_likelihood`` (sklearn
    ``_gpr.py:583-651``): Cholesky (failure -> (-inf, 0), ``:589-593``),
    alpha = K^-1 y, LML = -1/2 y'alpha - sum log L_ii - m/2 log 2pi
    (``:613-617``), gradient 1/2 tr((alpha alpha' - K^-1) dK/dtheta_j)
    (``:629-651``).  Returns (lml, grad[3], status) with status 0 ok / 1 not PD.
    """
    y = np.asarray(y, dtype=np.float64)
    K, dK = np_kernel(t, theta, eval_gradient=True)
    try:
        L = la.cholesky(K, lower=True, check_finite=False)
    except la.LinAlgError:
        return -np.inf, np.zeros(3), 1
    alpha = la.cho_solve((L, True), y, check_finite=False)
    lml = -0.5 * float(y @ alpha) - float(np.log(np.diag(L)).sum())
    lml -= 0.5 * K.shape[0] * LOG_2PI
    Kinv = la.cho_solve((L, True), np.eye(K.shape[0]), check_finite=False)
    inner = np.outer(alpha, alpha) - Kinv
    grad = 0.5 * np.einsum("ij,jik->k", inner, dK)
    return lml, grad, 0


def np_lml_grad_lean(t, y, theta, block=1024, want_alpha=False):
    """Same quantity as ``np_lml_grad`` (sklearn ``_gpr.py:583-651``, ``kernels.py:1559-1578``) for sizes where the
    reference's m x m x 3 gradient tensor does not fit: K, L and K^-1 = cho_solve(L, I) are formed exactly as the
    reference does (LAPACK dpotrf / dpotrs), the three traces 1/2 sum_ij (a_i a_j - K^-1_ij) dK_ij/dtheta_k are
    accumulated over row blocks with dK regenerated on the fly.  Returns (lml, grad[3], status[, alpha])."""
    y = np.asarray(y, dtype=np.float64)
    sig2, ell, chi = np.exp(np.asarray(theta, dtype=np.float64))
    x = np.asarray(t, dtype=np.float64) / ell
    m = x.size
    K = np.empty((m, m))
    for i0 in range(0, m, block):
        d = (x[i0:i0 + block, None] - x[None, :]) ** 2
        K[i0:i0 + block] = sig2 * np.exp(-0.5 * d)
    K[np.diag_indices(m)] = sig2 + chi              # R's diagonal is forced to 1 (kernels.py:1565)
    try:
        L = la.cholesky(K, lower=True, overwrite_a=True, check_finite=False)
    except la.LinAlgError:
        return (-np.inf, np.zeros(3), 1) + ((None,) if want_alpha else ())
    alpha = la.cho_solve((L, True), y, check_finite=False)
    lml = -0.5 * float(y @ alpha) - float(np.log(np.diag(L)).sum()) - 0.5 * m * LOG_2PI
    Kinv = la.cho_solve((L, True), np.eye(m), overwrite_b=True, check_finite=False)
    s = np.zeros(3)
    for i0 in range(0, m, block):
        i1 = min(m, i0 + block)
        d = (x[i0:i1, None] - x[None, :]) ** 2
        kr = sig2 * np.exp(-0.5 * d)
        kr[np.arange(i1 - i0), np.arange(i0, i1)] = sig2
        inner = np.outer(alpha[i0:i1], alpha) - Kinv[i0:i1]
        s[0] += float(np.sum(inner * kr))
        s[1] += float(np.sum(inner * (kr * d)))
    s[2] = chi * float(np.sum(alpha * alpha - np.diag(Kinv)))
    out = (lml, 0.5 * s, 0)
    return out + ((alpha,) if want_alpha else ())


def ld_truth(t, y, theta, exe=None):
    """Extended-precision (80-bit) LML / gradient / alpha from ``oracle/lml_ld.c`` (built by ``oracle/Makefile`` into
    ``oracle/_build/lml_ld``).  Returns (lml, grad[3], alpha[m])."""
    import os
    import struct
    import subprocess
    import tempfile

    exe = exe or os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "lml_ld")
    t = np.ascontiguousarray(t, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    m = t.size
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        with open(fin, "wb") as f:
            f.write(struct.pack("q", m))
            f.write(t.tobytes())
            f.write(y.tobytes())
            f.write(np.ascontiguousarray(theta, dtype=np.float64).tobytes())
        subprocess.run([exe, fin, fout], check=True)
        o = np.fromfile(fout)
    return float(o[0]), o[1:4].copy(), o[4:4 + m].copy()


def np_alpha(t, y, theta):
    """alpha_ = K^-1 y at fixed theta (sklearn ``_gpr.py:349-367``)."""
    K = np_kernel(t, theta)
    L = la.cholesky(K, lower=True, check_finite=False)
    return la.cho_solve((L, True), np.asarray(y, dtype=np.float64), check_finite=False), L


def np_predict(t, y, theta, t_star):
    """Posterior mean and standard deviation (sklearn ``_gpr.py:444-500``).

    mean = sigma^2 R(t*, t) alpha (no white-noise term in the cross kernel);
    var = (sigma^2 + chi) - sum_k V_ki^2, V = L^-1 k*', clipped at 0.
    """
    alpha, L = np_alpha(t, y, theta)
    Ks = np_kernel(t_star, theta, t2=t)
    mean = Ks @ alpha
    V = la.solve_triangular(L, Ks.T, lower=True, check_finite=False)
    sig2, _, chi = np.exp(np.asarray(theta, dtype=np.float64))
    var = (sig2 + chi) - np.einsum("ij,ij->j", V, V)
    var[var < 0] = 0.0
    return mean, np.sqrt(var)


def np_rbf_eval(t1, t2, sig2, ell):
    """kappa(t1, t2) = sigma^2 exp(-(t1 - t2)^2 / (2 ell^2)) (``gpkernels.py:591-609``)."""
    tdiff = np.asarray(t1)[:, None] - np.asarray(t2)
    return sig2 * np.exp(-(tdiff**2) / (2 * ell**2))


def np_lstsq_moments(t, y, theta, t_est, eta=1e-8, want_sqrtW=True):
    """state/derivative estimates, derivative covariance and sqrtW.

    Follows ``GP_RBFW.compute_lstsq_matrices`` (``gpkernels.py:612-649``) and
    ``_BaseGP._compute_estimates_and_weights`` (``gpkernels.py:445-504``).
    """
    sig2, ell, chi = np.exp(np.asarray(theta, dtype=np.float64))
    t = np.asarray(t, dtype=np.float64)
    t_est = np.asarray(t_est, dtype=np.float64)
    rbf_yy = np_rbf_eval(t, t, sig2, ell)
    rbf_zy = np_rbf_eval(t_est, t, sig2, ell)
    rbf_zz = np_rbf_eval(t_est, t_est, sig2, ell)
    dzz = t_est[:, None] - t_est
    dzy = t_est[:, None] - t
    ell2 = ell**2
    K_yy = rbf_yy + np.diag(np.full(t.size, chi))
    K_zy = -dzy * rbf_zy / ell2
    K_zz = (1 - (dzz**2 / ell2)) * rbf_zz / ell2

    cf = la.cho_factor(K_yy, check_finite=True)
    a = la.cho_solve(cf, np.asarray(y, dtype=np.float64))
    state = rbf_zy @ a
    ddt = K_zy @ a
    X = K_zy @ la.cho_solve(cf, K_zy.T)
    X = 0.5 * (X + X.T)
    C = K_zz - X
    out = dict(state_estimate=state, ddt_estimate=ddt, ddt_covariance=C)
    if want_sqrtW:
        out["sqrtW"] = np_sqrtW(C, eta)
    return out


def np_sqrtW(C, eta):
    """sqrtW = (C + eta I)^(-1/2) via eigh (``gpkernels.py:496-504``)."""
    ev, V = la.eigh(C + eta * np.eye(C.shape[0]), check_finite=False)
    if np.any(ev <= 0):
        raise ValueError("inverse covariance not positive definite, increase eta")
    return V @ np.diag(1 / np.sqrt(ev)) @ V.T


# --------------------------------------------------------------------------
# Matern extension (the reference has no Matern kernel; BASELINE.json's north
# star names it).  Oracle = scikit-learn's Matern for K / dK, the analytic
# derivatives of the same kernel for the cross-covariances.
# --------------------------------------------------------------------------
def np_matern(t1, t2, sig2, ell, twice_nu, deriv=0):
    """sigma^2 M_nu((t1 - t2) / ell) for nu = twice_nu / 2 in {3/2, 5/2} (sklearn
    ``kernels.py:1601-1790``: K = dists * sqrt(2 nu); (1 + K) exp(-K) resp.
    (1 + K + K^2/3) exp(-K)), or its derivatives w.r.t. the UNSCALED times:
    deriv=1: d k / d t1;  deriv=2: d^2 k / d t1 d t2 = -k''(tau)."""
    tau = np.asarray(t1, dtype=np.float64)[:, None] - np.asarray(t2, dtype=np.float64)[None, :]
    a = np.sqrt(float(twice_nu)) / ell
    K = a * np.abs(tau)
    e = np.exp(-K)
    if twice_nu == 3:
        if deriv == 0:
            return sig2 * (1.0 + K) * e
        if deriv == 1:
            return -sig2 * a**2 * tau * e
        return sig2 * a**2 * (1.0 - K) * e
    if twice_nu == 5:
        if deriv == 0:
            return sig2 * (1.0 + K + K**2 / 3.0) * e
        if deriv == 1:
            return -sig2 * a**2 / 3.0 * tau * (1.0 + K) * e
        return sig2 * a**2 / 3.0 * (1.0 + K - K**2) * e
    raise ValueError("twice_nu must be 3 or 5")


def sk_matern_kernel(theta, twice_nu, bounds=None):
    """scikit-learn's (ConstantKernel * Matern(nu)) + WhiteKernel at log-hyperparameters theta."""
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel

    s2, ell, chi = np.exp(np.asarray(theta, dtype=np.float64))
    b = [(1e-300, 1e300)] * 3 if bounds is None else [tuple(x) for x in bounds]
    return (ConstantKernel(s2, constant_value_bounds=b[0]) * Matern(length_scale=ell, nu=twice_nu / 2,
                                                                    length_scale_bounds=b[1])
            + WhiteKernel(chi, noise_level_bounds=b[2]))


def np_lml_grad_matern(t, y, theta, twice_nu):
    """LML and gradient exactly as sklearn's regressor computes them (``_gpr.py:583-651``) with the Matern kernel.
    Returns (lml, grad[3], status)."""
    y = np.asarray(y, dtype=np.float64)
    K, dK = sk_matern_kernel(theta, twice_nu)(np.asarray(t, dtype=np.float64)[:, None], eval_gradient=True)
    try:
        L = la.cholesky(K, lower=True, check_finite=False)
    except la.LinAlgError:
        return -np.inf, np.zeros(3), 1
    alpha = la.cho_solve((L, True), y, check_finite=False)
    lml = -0.5 * float(y @ alpha) - float(np.log(np.diag(L)).sum()) - 0.5 * K.shape[0] * LOG_2PI
    Kinv = la.cho_solve((L, True), np.eye(K.shape[0]), check_finite=False)
    grad = 0.5 * np.einsum("ij,jik->k", np.outer(alpha, alpha) - Kinv, dK)
    return lml, grad, 0


def np_predict_matern(t, y, theta, t_star, twice_nu):
    """Posterior mean / std with the Matern kernel (sklearn ``_gpr.py:444-500`` arithmetic)."""
    sig2, ell, chi = np.exp(np.asarray(theta, dtype=np.float64))
    K = np_matern(t, t, sig2, ell, twice_nu) + chi * np.eye(len(t))
    L = la.cholesky(K, lower=True, check_finite=False)
    alpha = la.cho_solve((L, True), np.asarray(y, dtype=np.float64), check_finite=False)
    Ks = np_matern(t_star, t, sig2, ell, twice_nu)
    V = la.solve_triangular(L, Ks.T, lower=True, check_finite=False)
    var = (sig2 + chi) - np.einsum("ij,ij->j", V, V)
    var[var < 0] = 0.0
    return Ks @ alpha, np.sqrt(var), alpha


def np_lstsq_moments_matern(t, y, theta, t_est, twice_nu):
    """The quantities of ``_compute_estimates_and_weights`` (``gpkernels.py:445-493``) with the Matern kernel:
    state = kappa_zy alpha, ddt = K_zy alpha, C = K_zz - K_zy K_yy^-1 K_zy^T, with K_zy = d kappa / d t',
    K_zz = d^2 kappa / d t' d t."""
    sig2, ell, chi = np.exp(np.asarray(theta, dtype=np.float64))
    K_yy = np_matern(t, t, sig2, ell, twice_nu) + chi * np.eye(len(t))
    k_zy = np_matern(t_est, t, sig2, ell, twice_nu)
    K_zy = np_matern(t_est, t, sig2, ell, twice_nu, deriv=1)
    K_zz = np_matern(t_est, t_est, sig2, ell, twice_nu, deriv=2)
    cf = la.cho_factor(K_yy, check_finite=True)
    a = la.cho_solve(cf, np.asarray(y, dtype=np.float64))
    X = K_zy @ la.cho_solve(cf, K_zy.T)
    return dict(state_estimate=k_zy @ a, ddt_estimate=K_zy @ a, ddt_covariance=K_zz - 0.5 * (X + X.T))


# --------------------------------------------------------------------------
# Restatement of gpkernels.GP_RBFW on top of scikit-learn (as the reference)
# --------------------------------------------------------------------------
class OracleGP:
    """Port of ``gpkernels.GP_RBFW`` (``gpkernels.py:507-649``) + ``_BaseGP``
    (``:299-504``): sklearn ``GaussianProcessRegressor(kernel=C*RBF+White,
    alpha=0, n_restarts_optimizer=...)``; restart points come from the global
    NumPy RNG exactly as in the reference (``_gpr.py:251,330``)."""

    def __init__(self, constant_bounds, length_scale_bounds, noise_level_bounds,
                 n_restarts_optimizer, twice_nu=0):
        from sklearn.gaussian_process import GaussianProcessRegressor
        from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern, WhiteKernel

        self.twice_nu = twice_nu      # 0: RBF (the reference); 3 / 5: Matern extension
        base = (RBF(length_scale_bounds=length_scale_bounds) if twice_nu == 0
                else Matern(length_scale_bounds=length_scale_bounds, nu=twice_nu / 2))
        kernel = (
            ConstantKernel(1.0, constant_value_bounds=constant_bounds) * base
        ) + WhiteKernel(noise_level_bounds=noise_level_bounds)
        self.gpr = GaussianProcessRegressor(
            kernel=kernel, n_restarts_optimizer=n_restarts_optimizer, alpha=0
        )

    # gpkernels.py:330-348
    def fit(self, t_training, training_data):
        if training_data.ndim > 1:
            raise ValueError("GP training data must be one-dimensional")
        self.t_training = t_training
        self.y = training_data
        self.gpr.fit(t_training[:, None], training_data)
        return self

    # gpkernels.py:350-365
    def predict(self, t):
        return self.gpr.predict(t[:, None], return_std=True)

    @property
    def theta(self):
        return np.array(self.gpr.kernel_.theta)

    @property
    def lml(self):
        return float(self.gpr.log_marginal_likelihood_value_)

    def lml_grad(self, theta):
        return self.gpr.log_marginal_likelihood(np.asarray(theta), eval_gradient=True)

    # gpkernels.py:612-649 / 445-504
    def compute_lstsq_matrices(self, t_est, eta=1e-8, want_sqrtW=True):
        out = np_lstsq_moments(self.t_training, self.y, self.theta, t_est, eta, want_sqrtW)
        self.t_estimation = t_est
        for k, v in out.items():
            setattr(self, k, v)
        return self


def oracle_fit_gaussian_processes(t_est, t_sampled, snapshots, bounds, n_restarts,
                                  eta=1e-8, want_sqrtW=True):
    """Port of ``fit_gaussian_processes`` (``PDEs/step2_fitgps.py:67-102``; a list of
    per-variable time vectors gives the ODE flavour, ``ODEs/step2_fitgps.py:68-97``)."""
    gps = []
    for i in range(snapshots.shape[0]):
        ti = t_sampled[i] if isinstance(t_sampled, (list, tuple)) else t_sampled
        gp = OracleGP(bounds[0], bounds[1], bounds[2], n_restarts)
        gp.fit(np.asarray(ti), np.asarray(snapshots[i]))
        gp.compute_lstsq_matrices(t_est, eta=eta, want_sqrtW=want_sqrtW)
        gps.append(gp)
    return gps


# --------------------------------------------------------------------------
# Synthetic workload of BASELINE.json's configs[3]/[4]
# --------------------------------------------------------------------------
def synthetic_trajectories(r, m, seed=0):
    """t = sort(U(0,1)) with endpoints forced, y_i = sum_k a sin(2 pi f t + phi) + 0.03 N(0,1)."""
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0.0, 1.0, size=m))
    t[0], t[-1] = 0.0, 1.0
    a = rng.uniform(0.5, 2.0, size=(r, 3))
    f = rng.uniform(1.0, 8.0, size=(r, 3))
    ph = rng.uniform(0.0, 2 * np.pi, size=(r, 3))
    y = (a[:, :, None] * np.sin(2 * np.pi * f[:, :, None] * t[None, None, :] + ph[:, :, None])).sum(1)
    y += 0.03 * rng.standard_normal((r, m))
    return t, y


def np_posterior_grid(sqrtW, D, rhs, regs):
    """CPU restatement of the step-3 posterior assembly for a grid of regularizers (PARITY UNPINNED:
    the reference's step 3 cannot run here because `opinf` (requirements.txt: opinf==0.5.9) is not installed).
    Follows PDEs/step3_estimate.py:75-95 with codebase/wlstsq.py:183-188:
      A_i = sqrtW_i @ D, b_i = sqrtW_i @ rhs_i                                  (wlstsq.py:183-188)
      mean = opinf.lstsq.L2Solver(reg).fit(A_i, b_i).solve(): the SVD route of opinf 0.5.9,
             V diag(s / (s^2 + reg^2)) U^T b                                   (published algorithm, un-vendored)
      precision = A_i^T A_i + reg^2 I                                           (step3_estimate.py:86-90)
      status = 1 where np.linalg.cholesky(precision) fails -- scipy.stats.Covariance.from_precision inside
               bayes.BayesianROM (bayes.py:283-287) raises "Matrix is not positive definite" there.
    -> means (K, r, d), chol (K, r, d, d), gram (r, d, d), proj (r, d), status (K, r)."""
    import numpy as np

    sqrtW, D, rhs = np.asarray(sqrtW, float), np.asarray(D, float), np.atleast_2d(np.asarray(rhs, float))
    regs = np.atleast_1d(np.asarray(regs, float))
    r, d, K = sqrtW.shape[0], D.shape[1], regs.size
    means, chol = np.empty((K, r, d)), np.zeros((K, r, d, d))
    gram, proj, status = np.empty((r, d, d)), np.empty((r, d)), np.zeros((K, r), dtype=np.int32)
    for i in range(r):
        A = sqrtW[i] @ D
        b = sqrtW[i] @ rhs[i]
        gram[i], proj[i] = A.T @ A, A.T @ b
        U, s, Vt = np.linalg.svd(A, full_matrices=False)
        Utb = U.T @ b
        for k, reg in enumerate(regs):
            means[k, i] = Vt.T @ ((s / (s**2 + reg**2)) * Utb)
            try:
                chol[k, i] = np.linalg.cholesky(gram[i] + reg**2 * np.eye(d))
            except np.linalg.LinAlgError:
                status[k, i] = 1
    return {"means": means, "chol": chol, "gram": gram, "proj": proj, "status": status}
