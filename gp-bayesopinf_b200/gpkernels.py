"""Drop-in replacement for the reference's ``codebase/gpkernels.py`` sklearn path (``GP_RBFW``).

Same constructor, methods, attributes, exceptions and ``str()`` format as the reference
(``gpkernels.py:299-649``); the arithmetic runs in ``libgpbo.so`` (hand-written sm_100a CUDA) through
the ctypes layer in ``_lib.py``.  Results surfaced to callers are host NumPy float64 arrays and the
object stays picklable (``save``/``load`` via joblib, ``gpkernels.py:423-430``).

Not re-implemented here (out of the hot path): the float32 gpytorch variant
``TORCH_GP_RBFW`` (``gpkernels.py:32-297``).

``sqrtW = (C + eta I)^(-1/2)`` (``gpkernels.py:496-504``) is computed on the GPU as well, while the covariance is
still in HBM, by a coupled Newton-Schulz iteration on the FP64 tensor tiles (``csrc/kernels_sqrtw.cuh``) instead of
the reference's ``eigh``; it is ill-posed element-wise and is checked through the identity
``sqrtW (C + eta I) sqrtW = I``.
"""

from __future__ import annotations

import warnings

import numpy as np

from . import _lib

__all__ = ["GP_RBFW", "GP_MaternW", "ConvergenceWarning"]


def _convergence_warning_class():
    """The warning CLASS scikit-learn raises (``_gpr.py:338, 667``) when scikit-learn is installed -- users of the
    reference filter on ``sklearn.exceptions.ConvergenceWarning`` -- else a local class of the same name.  Only the
    exception module is touched: no scikit-learn / SciPy numerics are ever used by this package."""
    import importlib

    try:
        return importlib.import_module("sklearn.exceptions").ConvergenceWarning
    except Exception:
        class ConvergenceWarning(UserWarning):
            """Optimum close to a bound / optimiser did not converge."""

        return ConvergenceWarning


ConvergenceWarning = _convergence_warning_class()


# ---- tiny stand-ins for the sklearn objects whose attributes the reference reads --------------
class _Const:
    def __init__(self, v):
        self.constant_value = v


class _Rbf:
    def __init__(self, v):
        self.length_scale = v


class _White:
    def __init__(self, v):
        self.noise_level = v


class _Prod:
    def __init__(self, k1, k2):
        self.k1, self.k2 = k1, k2


class _KernelState:
    """Mirror of sklearn's fitted ``kernel_`` for (C * RBF) + White: ``theta``, ``bounds``, ``k1.k1`` ..."""

    def __init__(self, theta, bounds_log):
        self.bounds = np.array(bounds_log, dtype=np.float64)
        self.theta = theta

    @property
    def theta(self):
        return self._theta

    @theta.setter
    def theta(self, th):
        th = np.array(th, dtype=np.float64)
        self._theta = th
        s2, ell, chi = np.exp(th)
        self.k1 = _Prod(_Const(float(s2)), _Rbf(float(ell)))
        self.k2 = _White(float(chi))

    def __repr__(self):
        return (f"{np.sqrt(self.k1.k1.constant_value):.3g}**2 * RBF(length_scale={self.k1.k2.length_scale:.3g})"
                f" + WhiteKernel(noise_level={self.k2.noise_level:.3g})")


class _GPRState:
    """What the reference exposes through ``gp.gpr`` after ``fit`` (sklearn attribute names)."""

    def __init__(self, bounds_log, n_restarts_optimizer):
        self.n_restarts_optimizer = int(n_restarts_optimizer)
        self.alpha = 0
        self.bounds_log = np.array(bounds_log, dtype=np.float64)
        self.kernel_ = None
        self._ctx_device = None
        self._twice_nu = 0

    def log_marginal_likelihood(self, theta=None, eval_gradient=False):
        """GPU evaluation of sklearn's ``log_marginal_likelihood`` (_gpr.py:541-656)."""
        if theta is None:
            return self.log_marginal_likelihood_value_
        ctx = _lib.default_context(self._ctx_device)
        ctx.set_kernel_family(self._twice_nu)
        lml, grad, _ = ctx.lml_grad(self.X_train_[:, 0][None, :], self.y_train_[None, :],
                                    np.asarray(theta, dtype=np.float64)[None, :])
        if eval_gradient:
            return float(lml[0]), grad[0]
        return float(lml[0])


def _check_bounds(theta, bounds_log, names=("k1__k1__constant_value", "k1__k2__length_scale", "k2__noise_level")):
    """sklearn ``kernels.py:436-465`` (``_check_bounds_params``): warn when the optimum sits on a bound.  Like
    sklearn, the comparison is ``np.isclose`` on the LOG values (theta against the log-bounds)."""
    theta = np.asarray(theta, dtype=np.float64)
    for i, name in enumerate(names):
        if np.isclose(theta[i], bounds_log[i, 0]):
            warnings.warn(f"The optimal value found for dimension 0 of parameter {name} is close to the specified "
                          f"lower bound {np.exp(bounds_log[i, 0])}. Decreasing the bound and calling fit again may "
                          "find a better value.", ConvergenceWarning)
        if np.isclose(theta[i], bounds_log[i, 1]):
            warnings.warn(f"The optimal value found for dimension 0 of parameter {name} is close to the specified "
                          f"upper bound {np.exp(bounds_log[i, 1])}. Increasing the bound and calling fit again may "
                          "find a better value.", ConvergenceWarning)


_LBFGS_STATUS = {2: "ABNORMAL_TERMINATION_IN_LNSRCH", 3: "STOP: TOTAL NO. of ITERATIONS REACHED LIMIT",
                 4: "STOP: TOTAL NO. of f AND g EVALUATIONS EXCEEDS LIMIT"}


def _warn_unconverged_starts(statuses):
    """sklearn checks EVERY start's optimiser result (``_gpr.py:667`` -> ``_check_optimize_result``) and warns for
    each one that did not converge, not only for the best."""
    for st in np.asarray(statuses).ravel():
        msg = _LBFGS_STATUS.get(int(st))
        if msg is not None:
            warnings.warn(f"lbfgs failed to converge (status={int(st)}):\n{msg}.\n\nIncrease the number of iterations "
                          "(max_iter) or scale the data as shown in:\n"
                          "    https://scikit-learn.org/stable/modules/preprocessing.html", ConvergenceWarning)


def draw_restart_points(bounds_log, n_restarts):
    """Restart points exactly as sklearn draws them (_gpr.py:251, 330): ``n_restarts`` calls of
    ``uniform(log lo, log hi)`` on the GLOBAL NumPy RandomState (SURVEY.md §8b RNG contract)."""
    rng = np.random.mtrand._rand
    b = np.asarray(bounds_log, dtype=np.float64)
    return np.array([rng.uniform(b[:, 0], b[:, 1]) for _ in range(int(n_restarts))]).reshape(-1, 3)


class GP_RBFW:
    """Gaussian process regressor with kernel  sigma^2 exp(-(t-t')^2 / (2 ell^2)) + chi delta(t,t').

    Mirrors ``gpkernels.GP_RBFW`` (``gpkernels.py:507-649``) and ``_BaseGP`` (``:299-504``).
    """

    def __init__(self, constant_bounds=(1e-5, 1e5), length_scale_bounds=(1.5e-6, 0.002),
                 noise_level_bounds=(1e-14, 1e-10), n_restarts_optimizer=50):
        bounds = np.array([constant_bounds, length_scale_bounds, noise_level_bounds], dtype=np.float64)
        self.gpr = _GPRState(np.log(bounds), n_restarts_optimizer)

    _twice_nu = 0          # kernel family of the library calls: 0 = RBF (the reference's kernel)

    def _context(self):
        ctx = _lib.default_context()
        ctx.set_kernel_family(self._twice_nu)
        return ctx

    # Properties ----------------------------------------------------------------
    @property
    def nsamples(self):
        if hasattr(self, "train_indices"):
            return self.t_training.size

    @property
    def constant(self):
        return self.gpr.kernel_.k1.k1.constant_value

    @property
    def length_scale(self):
        return self.gpr.kernel_.k1.k2.length_scale

    @property
    def noise_level(self):
        return self.gpr.kernel_.k2.noise_level

    def __str__(self):
        return "\n\t".join([
            "Gaussian radial basis function kernel",
            r"k(t, t') = \sigma^2 exp(-(t - t')^2 / (2 \ell^2)) + \chi I",
            rf"\sigma^2 = {self.constant:.4e}",
            rf"\ell = {self.length_scale:.4e}",
            rf"\chi = {self.noise_level:.4e}",
        ])

    # Main routines -----------------------------------------------------------------
    def fit(self, t_training, training_data):
        """Multi-restart maximisation of the log-marginal likelihood (gpkernels.py:330-348 ->
        sklearn _gpr.py:233-368): start 0 at theta = log(1, 1, 1), then ``n_restarts_optimizer``
        starts drawn from the global NumPy RNG; all starts advance in lock-step on the GPU."""
        training_data = np.asarray(training_data)
        if training_data.ndim > 1:
            raise ValueError("GP training data must be one-dimensional")
        t_training = np.asarray(t_training, dtype=np.float64)
        if not (np.all(np.isfinite(t_training)) and np.all(np.isfinite(np.asarray(training_data, dtype=np.float64)))):
            raise ValueError("Input contains NaN or infinity.")      # sklearn validate_data in GaussianProcessRegressor.fit
        starts = np.vstack([np.zeros((1, 3)), draw_restart_points(self.gpr.bounds_log, self.gpr.n_restarts_optimizer)])
        ctx = self._context()
        res = ctx.fit(t_training[None, :], training_data[None, :], self.gpr.bounds_log, starts,
                      gp_of=np.zeros(len(starts), dtype=np.int32))
        self._set_fit_result(t_training, training_data, res["theta"], res["fun"], res["status"])
        self._finish_fit(ctx)
        return self

    def _set_fit_result(self, t_training, training_data, thetas, funs, statuses):
        self.t_training = t_training
        self.y = training_data
        funs = np.where(np.isnan(funs), np.inf, funs)
        best = int(np.argmin(funs))                      # _gpr.py:336-337
        theta = thetas[best]
        self.gpr.kernel_ = _KernelState(theta, self.gpr.bounds_log)
        _check_bounds(theta, self.gpr.bounds_log)        # _gpr.py:338
        self.gpr.log_marginal_likelihood_value_ = -float(funs[best])
        self.gpr.X_train_ = np.array(t_training, dtype=np.float64)[:, None]
        self.gpr.y_train_ = np.array(training_data, dtype=np.float64)
        _warn_unconverged_starts(statuses)               # _gpr.py:667, once per start

    def _finish_fit(self, ctx, alpha=None, status=None):
        """alpha_ = K^-1 y at the selected theta (_gpr.py:349-367); non-PD -> LinAlgError."""
        if alpha is None:
            _, _, alpha, status = ctx.predict(self.t_training[None, :], self.y[None, :], self.gpr.kernel_.theta[None, :],
                                              self.t_training[:1], want_alpha=True)
            alpha, status = alpha[0], status[0]
        if status != 0:
            raise np.linalg.LinAlgError(
                f"The kernel, {self.gpr.kernel_}, is not returning a positive definite matrix. Try gradually "
                "increasing the 'alpha' parameter of your GaussianProcessRegressor estimator.")
        self.gpr.alpha_ = alpha

    def predict(self, t):
        """(mean, std) of the posterior at ``t`` (gpkernels.py:350-365 -> _gpr.py:444-500)."""
        t = np.asarray(t, dtype=np.float64)
        ctx = self._context()
        mean, std, _, st = ctx.predict(self.t_training[None, :], self.y[None, :], self.gpr.kernel_.theta[None, :], t)
        if st[0] != 0:
            raise np.linalg.LinAlgError("kernel matrix not positive definite")
        if np.any(std[0] == 0.0):
            warnings.warn("Predicted variances smaller than 0. Setting those variances to 0.")
        return mean[0], std[0]

    def prediction_bounds(self, t, kind="95%"):
        mean, std = self.predict(t)
        if kind == "std":
            width = std
        elif kind == "95%":
            width = 1.96 * std
        elif kind == "2std":
            width = 2 * std
        elif kind == "3std":
            width = 3 * std
        else:
            raise ValueError(kind)
        return mean - width, mean, mean + width

    def _assemble(self, kind, t1, t2):
        import torch  # tensor hand-off only

        ctx = self._context()
        dev = torch.device("cuda", ctx.device)
        a = torch.as_tensor(np.ascontiguousarray(t1, dtype=np.float64), device=dev)
        b = torch.as_tensor(np.ascontiguousarray(t2, dtype=np.float64), device=dev)
        th = torch.as_tensor(self.gpr.kernel_.theta, device=dev)
        out = torch.empty((a.numel(), b.numel()), dtype=torch.float64, device=dev)
        torch.cuda.synchronize(dev)
        ctx.assemble_device(kind, a.data_ptr(), 0, a.numel(), b.data_ptr(), 0, b.numel(), th.data_ptr(), 1,
                            out.data_ptr(), 0, twice_nu=self._twice_nu)
        return out.cpu().numpy()

    def __call__(self, t, tprime):
        """kernel_(t, t') = sigma^2 R(t, t') (no white-noise term; gpkernels.py:405-420)."""
        return self._assemble(2, t, tprime)

    def rbf_eval(self, t1, t2):
        """kappa(t1, t2) (gpkernels.py:591-609)."""
        return self._assemble(3, t1, t2)

    # Persistence -----------------------------------------------------------------
    def save(self, save_path):
        import joblib

        joblib.dump(self, save_path)

    @staticmethod
    def load(load_path):
        import joblib

        return joblib.load(load_path)

    # Least-squares data ---------------------------------------------------------------
    def compute_lstsq_matrices(self, t_est, eta=1e-8):
        """state_estimate, ddt_estimate, ddt_covariance, sqrtW at ``t_est`` (gpkernels.py:612-649, 445-504)."""
        t_est = np.asarray(t_est, dtype=np.float64)
        ctx = self._context()
        state, ddt, cov, w, st, wst, _ = ctx.lstsq_weights(self.t_training[None, :], self.y[None, :],
                                                           self.gpr.kernel_.theta[None, :], t_est, eta)
        self._set_lstsq_result(t_est, state[0], ddt[0], cov[0], int(st[0]), w[0], int(wst[0]))
        return None

    def _set_lstsq_result(self, t_est, state, ddt, cov, status, sqrtW, w_status, with_sqrtW=True):
        """``cov`` / ``sqrtW`` may be None on a rank that does not hold them (multi-GPU, ``gather_cov=False``): the
        attributes are then None; the status checks are the same on every rank."""
        self.t_estimation = t_est
        if status != 0 or (cov is not None and not np.all(np.isfinite(cov))):
            # scipy.linalg.cho_factor(K_yy, check_finite=True) raises for the same inputs (gpkernels.py:481)
            raise np.linalg.LinAlgError("K_yy is not positive definite")
        self.state_estimate = state
        self.ddt_estimate = ddt
        self.ddt_covariance = cov
        if not with_sqrtW:
            return
        if w_status != 0:                                  # gpkernels.py:500-503
            raise ValueError("inverse covariance not positive definite, increase eta")
        self.sqrtW = sqrtW


class GP_MaternW(GP_RBFW):
    """Same interface with a Matern kernel, ``sigma^2 M_nu(|t - t'| / ell) + chi delta(t, t')``, nu in {1.5, 2.5}.

    An extension: the reference has only the RBF kernel.  Semantics follow scikit-learn's
    ``ConstantKernel * Matern(nu) + WhiteKernel`` (``sklearn/gaussian_process/kernels.py:1601-1790``) for the fit,
    the LML and ``predict``; ``compute_lstsq_matrices`` uses the analytic derivatives of the Matern kernel
    (``K_zy = d kappa / d t'``, ``K_zz = d^2 kappa / d t' d t``) in the reference's formulas (``gpkernels.py:445-504``).
    ``rbf_eval`` returns the Matern kernel value (no noise term), like ``__call__``.
    """

    def __init__(self, nu=2.5, constant_bounds=(1e-5, 1e5), length_scale_bounds=(1e-5, 1e2),
                 noise_level_bounds=(1e-16, 1e2), n_restarts_optimizer=50):
        if nu not in (1.5, 2.5):
            raise ValueError("GP_MaternW supports nu = 1.5 and nu = 2.5")
        super().__init__(constant_bounds, length_scale_bounds, noise_level_bounds, n_restarts_optimizer)
        self.nu = nu
        self._twice_nu = int(round(2 * nu))
        self.gpr._twice_nu = self._twice_nu

    def __str__(self):
        return "\n\t".join([
            f"Gaussian process with Matern kernel, nu = {self.nu}",
            r"k(t, t') = \sigma^2 M_\nu(|t - t'| / \ell) + \chi I",
            rf"\sigma^2 = {self.constant:.4e}",
            rf"\ell = {self.length_scale:.4e}",
            rf"\chi = {self.noise_level:.4e}",
        ])
