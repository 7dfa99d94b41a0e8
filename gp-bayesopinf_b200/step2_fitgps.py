"""Batched drop-in for the reference's ``*/step2_fitgps.py``.

``fit_gaussian_processes`` keeps the reference signatures
(``PDEs/step2_fitgps.py:67-72``: one shared sample-time vector; ``ODEs/step2_fitgps.py:68-73``: a list of
per-variable sample-time vectors) and return type (``list[GP_RBFW]``), but instead of looping over the
modes sequentially it

1. draws every GP's restart points from the global NumPy RNG in the reference's order (GP by GP,
   ``n_restarts`` draws each -- SURVEY.md §8b RNG contract),
2. optimises ALL (mode x start) pairs in lock-step on the GPU (one batched launch per line-search
   probe), optionally sharded over the ranks of a ``torch.distributed`` process group,
3. selects each GP's best start (``np.argmin``, sklearn ``_gpr.py:336-337``),
4. evaluates the posterior moments of all GPs in one batched pass.

Hyper-parameter bounds / restart count come from the experiment's ``config`` module exactly as in the
reference (``config.CONSTANT_VALUE_BOUNDS`` ...), or from the keyword overrides.
"""

from __future__ import annotations

import numpy as np

from . import _lib, sharding
from .gpkernels import GP_MaternW, GP_RBFW, draw_restart_points

__all__ = ["fit_gaussian_processes", "fit_gaussian_processes_multi"]


def _config_value(name, override):
    if override is not None:
        return override
    try:
        import config  # the experiment directory's config.py, as in the reference
    except Exception as exc:  # pragma: no cover
        raise RuntimeError(f"`config.{name}` is needed (or pass it as a keyword argument)") from exc
    return getattr(config, name)


def _as_time_matrix(time_domain_sampled, num_vars, sample_size):
    if isinstance(time_domain_sampled, (list, tuple)):
        if len(time_domain_sampled) != num_vars:
            raise ValueError("time_domains_sampled and snapshots_sampled not aligned")
        T = np.array([np.asarray(t, dtype=np.float64) for t in time_domain_sampled])
    else:
        t = np.asarray(time_domain_sampled, dtype=np.float64)
        if t.ndim == 2:
            T = t
        else:
            if t.size != sample_size:
                raise ValueError("time_domain_sampled and snapshots_sampled not aligned")
            T = np.tile(t, (num_vars, 1))
    if T.shape != (num_vars, sample_size):
        raise ValueError("time_domain_sampled and snapshots_sampled not aligned")
    return np.ascontiguousarray(T)


def fit_gaussian_processes(time_domain_training, time_domain_sampled, snapshots_sampled, gp_regularizer=1e-8, *,
                           constant_bounds=None, length_scale_bounds=None, noise_level_bounds=None,
                           n_restarts_optimizer=None, verbose=True, group=None, want_sqrtW=True, kernel="rbf",
                           gather_cov=True):
    """Fit one GP per row of ``snapshots_sampled`` and compute its least-squares data.

    Parameters follow the reference; ``time_domain_sampled`` may be one (m,) vector (PDE flavour) or a
    list of per-variable (m,) vectors (ODE flavour).  Returns ``list[GP_RBFW]``.

    Multi-GPU (``group=True`` or a ``torch.distributed`` group): see ``fit_gaussian_processes_multi``.
    """
    gps = fit_gaussian_processes_multi(
        time_domain_training, [time_domain_sampled], [snapshots_sampled], gp_regularizer,
        constant_bounds=constant_bounds, length_scale_bounds=length_scale_bounds,
        noise_level_bounds=noise_level_bounds, n_restarts_optimizer=n_restarts_optimizer, verbose=verbose,
        group=group, want_sqrtW=want_sqrtW, kernel=kernel, gather_cov=gather_cov)
    return gps[0]


_KERNELS = {"rbf": 0, "matern32": 3, "matern52": 5}


def fit_gaussian_processes_multi(time_domain_training, time_domains_sampled, snapshots_list, gp_regularizer=1e-8, *,
                                 constant_bounds=None, length_scale_bounds=None, noise_level_bounds=None,
                                 n_restarts_optimizer=None, verbose=True, group=None, want_sqrtW=True, kernel="rbf",
                                 gather_cov=True):
    """Multi-trajectory form: the loop of ``PDEsMulti/main.py:99-109`` as ONE batch (L trajectories x r modes).

    ``time_domains_sampled[l]`` / ``snapshots_list[l]`` are trajectory l's sample times and (r, m) data.
    Returns ``gps[l][i]``.  All trajectories must share the sample count m (they do in the reference).

    With a process group the evaluations of the optimiser are spread over the ranks (``sharding.fit_pairs``) and every
    rank returns the same fitted GPs.  ``gather_cov=True`` (default) also gives every rank every GP's
    ``ddt_covariance`` / ``sqrtW``, which the reference's step 3 reads for all modes (``PDEs/step3_estimate.py:212``,
    ``PDEs/main.py:219``) -- so the rest of ``main.py`` runs unchanged on every rank.  ``gather_cov=False`` keeps the
    two m' x m' matrices on the rank that computed them (GP g on rank ``g % world``) and sets them to ``None``
    elsewhere: step 3 must then run where the matrices are.  Numerical failures (non-PD ``K_yy``, indefinite
    ``C + eta I``) are all-gathered, so every rank raises the same exception.
    """
    if kernel not in _KERNELS:
        raise ValueError(f"kernel must be one of {sorted(_KERNELS)}")    # "rbf" is the reference's kernel
    twice_nu = _KERNELS[kernel]
    cb = _config_value("CONSTANT_VALUE_BOUNDS", constant_bounds)
    lb = _config_value("LENGTH_SCALE_BOUNDS", length_scale_bounds)
    nb = _config_value("NOISE_LEVEL_BOUNDS", noise_level_bounds)
    nres = int(_config_value("N_RESTARTS_OPTIMIZER", n_restarts_optimizer))
    t_est = np.ascontiguousarray(time_domain_training, dtype=np.float64)

    Ts, Ys, owner = [], [], []
    for ell, (tt, Q) in enumerate(zip(time_domains_sampled, snapshots_list)):
        Q = np.asarray(Q, dtype=np.float64)
        if Q.ndim != 2:
            raise ValueError("snapshots_sampled must be two-dimensional")
        num_vars, sample_size = Q.shape
        Ts.append(_as_time_matrix(tt, num_vars, sample_size))
        Ys.append(Q)
        owner += [ell] * num_vars
    sizes = {T.shape[1] for T in Ts}
    if len(sizes) != 1:
        raise ValueError("all trajectories must have the same number of samples")
    T = np.ascontiguousarray(np.vstack(Ts))
    Y = np.ascontiguousarray(np.vstack(Ys))
    G = T.shape[0]
    if not (np.all(np.isfinite(T)) and np.all(np.isfinite(Y))):
        raise ValueError("Input contains NaN or infinity.")          # sklearn validate_data in GaussianProcessRegressor.fit

    # 1. restart points, in the reference's order (all ranks draw the same stream)
    objs = [GP_RBFW(cb, lb, nb, nres) if twice_nu == 0 else GP_MaternW(twice_nu / 2, cb, lb, nb, nres) for _ in range(G)]
    bounds_log = objs[0].gpr.bounds_log
    S = nres + 1
    starts = np.zeros((G, S, 3))
    for g in range(G):
        starts[g, 1:] = draw_restart_points(bounds_log, nres)
    gp_of = np.repeat(np.arange(G, dtype=np.int32), S)

    # 2. lock-step optimisation of all pairs (sharded over ranks when a process group is given)
    ctx = _lib.default_context()
    ctx.set_kernel_family(twice_nu)
    res = sharding.fit_pairs(ctx, T, Y, bounds_log, starts.reshape(-1, 3), gp_of, group=group)
    thetas = res["theta"].reshape(G, S, 3)
    funs = res["fun"].reshape(G, S)
    stats = res["status"].reshape(G, S)

    # 3. best start per GP
    for g in range(G):
        objs[g]._set_fit_result(T[g], Y[g], thetas[g], funs[g], stats[g])
    theta_opt = np.array([o.gpr.kernel_.theta for o in objs])

    # 4. posterior moments of every GP in one batched pass
    mom = sharding.moments(ctx, T, Y, theta_opt, t_est, group=group, eta=gp_regularizer if want_sqrtW else None,
                           gather_cov=gather_cov)
    for g in range(G):
        o = objs[g]
        o._finish_fit(ctx, alpha=mom["alpha"][g], status=int(mom["fit_status"][g]))
        if verbose:
            print(o)
        # same checks, same exceptions on every rank (the statuses were all-gathered), matrices where they exist
        o._set_lstsq_result(t_est, mom["state"][g], mom["ddt"][g], mom["cov"][g], int(mom["status"][g]),
                            mom["sqrtW"][g] if want_sqrtW else None, int(mom["w_status"][g]) if want_sqrtW else 0,
                            with_sqrtW=want_sqrtW)

    out, k = [], 0
    for Q in Ys:
        out.append(objs[k:k + Q.shape[0]])
        k += Q.shape[0]
    return out

