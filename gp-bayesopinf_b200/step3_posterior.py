"""Posterior assembly of step 3 on the GPU -- the linear algebra of ``PDEs/step3_estimate.py:75-95``
(``get_bayesian_model(reg)``: posterior mean of every operator row, precision ``(sqrtW_i D)^T (sqrtW_i D) + reg^2 I``)
evaluated for a whole grid of regularizers in one call (the grid search of ``:131-148`` tries 81 of them, each a
separate solve on the CPU).

What this module does NOT do is the rest of the grid search: for every candidate the reference draws operator samples
and integrates the reduced-order model through ``opinf`` (``_training_error``, ``:102-129``); that needs the ``opinf``
package and stays where it is.  The outputs here are exactly what that code consumes: ``means`` (what
``rom._extract_operators`` takes), ``precisions`` (what ``bayes.BayesianROM`` takes) and -- to skip the Cholesky
``scipy.stats.Covariance.from_precision`` repeats per candidate (``codebase/bayes.py:283-287``) -- the lower Cholesky
factor of every precision, plus the reference's "Matrix is not positive definite" outcome as a flag.
"""
from __future__ import annotations

import numpy as np

from . import _lib


class PosteriorGrid:
    """Result of :func:`posterior_grid`; ``k`` indexes the regularizer, ``i`` the mode / GP."""

    def __init__(self, regularizers, res):
        self.regularizers = np.asarray(regularizers, dtype=np.float64)
        self.means = res["means"]              # (K, r, d)   lstsq_solver.solve() per candidate
        self.chol = res["chol"]                # (K, r, d, d) lower Cholesky factors of the precisions (or None)
        self.gram = res["gram"]                # (r, d, d)   (sqrtW_i D)^T (sqrtW_i D)
        self.proj = res["proj"]                # (r, d)      (sqrtW_i D)^T (sqrtW_i z_i)
        self.status = res["status"]            # (K, r)      1: precision not positive definite

    def precisions(self, k):
        """The ``precisions`` list of ``get_bayesian_model`` for candidate k (``step3_estimate.py:84-90``)."""
        reg2 = float(self.regularizers[k]) ** 2
        eye = np.eye(self.gram.shape[1])
        return [g + reg2 * eye for g in self.gram]

    def is_spd(self, k):
        """False where the reference's ``BayesianROM`` construction raises "Matrix is not positive definite" and
        ``get_bayesian_model`` returns None (``step3_estimate.py:91-96``)."""
        return bool(np.all(self.status[k] == 0))

    def draw(self, k, rng=None):
        """One posterior draw of the operator matrix for candidate k, the way ``scipy.stats.multivariate_normal`` with
        a ``Covariance.from_precision`` object colours white noise: ``mean + C^-T z`` (``bayes.py:332-335``)."""
        if self.chol is None:
            raise ValueError("posterior_grid was called with want_chol=False")
        if not self.is_spd(k):
            raise np.linalg.LinAlgError("Matrix is not positive definite")
        from_rng = np.random if rng is None else rng
        out = np.empty_like(self.means[k])
        for i in range(out.shape[0]):
            z = from_rng.standard_normal(out.shape[1])
            out[i] = self.means[k, i] + np.linalg.solve(self.chol[k, i].T, z)   # upper-triangular d x d, d <= 128
        return out


def posterior_grid(gps_or_sqrtW, data_matrix, rhs, regularizers, *, ctx=None, want_chol=True):
    """Posterior means / precisions of all operator rows for every regularizer of ``regularizers``.

    gps_or_sqrtW : list of fitted ``GP_RBFW`` (their ``sqrtW`` is used, as ``estimate_posterior`` does,
        ``step3_estimate.py:209-212``), an ``(r, m', m')`` array of weight matrices, or ``None`` to use the matrices still
        resident in HBM from the ``compute_lstsq_matrices`` / ``fit_gaussian_processes`` call that produced them.
    data_matrix : ``(m', d)`` the unweighted data matrix D (``rom._assemble_data_matrix``), d <= 128.
    rhs : ``(r, m')`` time-derivative estimates (``gp.ddt_estimate``).
    """
    ctx = ctx or _lib.default_context()
    w = None
    if gps_or_sqrtW is not None:
        if isinstance(gps_or_sqrtW, (list, tuple)) and hasattr(gps_or_sqrtW[0], "sqrtW"):
            if any(getattr(gp, "sqrtW", None) is None for gp in gps_or_sqrtW):
                raise ValueError("a GP has no sqrtW (fit with want_sqrtW=True, or gather_cov=True on every rank)")
            w = np.array([gp.sqrtW for gp in gps_or_sqrtW])
        else:
            w = np.asarray(gps_or_sqrtW, dtype=np.float64)
            if w.ndim == 2:
                w = w[None]
    regs = np.atleast_1d(np.asarray(regularizers, dtype=np.float64))
    return PosteriorGrid(regs, ctx.posterior_grid(data_matrix, rhs, regs, sqrtW=w, want_chol=want_chol))
