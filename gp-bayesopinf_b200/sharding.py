"""Multi-GPU sharding of the step2_fitgps batch.

The reference is single-process; here the (GP x start) pairs are independent optimisation problems
and the GPs are independent for prediction, so one process per GPU takes a cyclic slice of the pairs
(cyclic spreads optimiser stragglers), and the only exchanges are

* one all-gather of ``[theta_opt(3), fun, nfev, nit, status]`` per pair (56 B), after which every
  rank performs the per-GP argmin of sklearn ``_gpr.py:336-340``;
* one all-gather of the per-GP posterior moments ``alpha_, state_estimate, ddt_estimate`` (+ status).

``ddt_covariance`` and ``sqrtW`` (m'^2 per GP each) stay on the owning rank (``cov[g] is None`` elsewhere).
The collectives go through ``torch.distributed`` (NCCL over NVLink on GPUs; gloo in the CPU tests).
``engine`` is anything with the ``_lib.Context`` methods ``fit`` / ``lstsq_moments`` / ``predict``.
"""

from __future__ import annotations

import numpy as np


def _dist(group):
    if group is None:
        return None, 0, 1
    import torch.distributed as dist

    if group is True:
        group = dist.group.WORLD
    return dist, dist.get_rank(group), dist.get_world_size(group)


def shard_indices(n, rank, world):
    """Cyclic shard: items rank, rank+world, ..."""
    return np.arange(rank, n, world)


def _all_gather_rows(dist, group, local, n_total, rank, world):
    """All-gather row blocks of a cyclic partition back into global order.  local: (n_local, k) float64."""
    import torch

    g = None if group is True else group
    k = local.shape[1]
    per = (n_total + world - 1) // world
    backend = dist.get_backend(g)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    buf = torch.zeros((per, k), dtype=torch.float64)
    buf[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local))
    buf = buf.to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=g)
    full = np.empty((n_total, k))
    for r in range(world):
        idx = shard_indices(n_total, r, world)
        full[idx] = out[r].cpu().numpy()[: idx.size]
    return full


def fit_pairs(engine, T, Y, bounds_log, starts, gp_of, group=None, opts=None):
    """Optimise all pairs; with a process group each rank optimises pairs ``rank::world``."""
    dist, rank, world = _dist(group)
    B = starts.shape[0]
    if world == 1:
        return engine.fit(T, Y, bounds_log, starts, gp_of=gp_of, opts=opts)
    idx = shard_indices(B, rank, world)
    if idx.size:
        loc = engine.fit(T, Y, bounds_log, starts[idx], gp_of=gp_of[idx], opts=opts)
        packed = np.column_stack([loc["theta"], loc["fun"], loc["nfev"], loc["nit"], loc["status"]])
        evals, rounds = loc["evals"], loc["rounds"]
    else:
        packed = np.zeros((0, 7))
        evals, rounds = 0, 0
    full = _all_gather_rows(dist, group, packed, B, rank, world)
    return dict(theta=full[:, :3].copy(), fun=full[:, 3].copy(), nfev=full[:, 4].astype(np.int32),
                nit=full[:, 5].astype(np.int32), status=full[:, 6].astype(np.int32), evals=evals, rounds=rounds)


def moments(engine, T, Y, theta_opt, t_est, group=None, want_cov=True, eta=None):
    """alpha_, state/ddt estimates, derivative covariance (and, when ``eta`` is given, sqrtW) of every GP;
    GPs ``rank::world`` per rank.  ``cov[g]`` / ``sqrtW[g]`` are None on ranks that do not own GP g."""
    dist, rank, world = _dist(group)
    G, m = T.shape
    n = t_est.shape[-1]
    idx = shard_indices(G, rank, world)
    cov = [None] * G
    sqrtW = [None] * G
    w_status = np.zeros(G, dtype=np.int32)
    if idx.size:
        pts = t_est if t_est.ndim == 1 else t_est[idx]
        if eta is not None:
            state, ddt, c, w, st, wst, _ = engine.lstsq_weights(T[idx], Y[idx], theta_opt[idx], pts, eta)
        else:
            state, ddt, c, st = engine.lstsq_moments(T[idx], Y[idx], theta_opt[idx], pts, want_cov=want_cov)
            w, wst = None, None
        _, _, alpha, fst = engine.predict(T[idx], Y[idx], theta_opt[idx], T[idx][:, :1], want_alpha=True)
        for k, g in enumerate(idx):
            cov[g] = c[k] if c is not None else None
            if w is not None:
                sqrtW[g] = w[k]
                w_status[g] = wst[k]
        packed = np.column_stack([alpha, state, ddt, st, fst])
    else:
        packed = np.zeros((0, m + 2 * n + 2))
    if world == 1:
        full = packed
    else:
        full = _all_gather_rows(dist, group, packed, G, rank, world)
    return dict(alpha=full[:, :m].copy(), state=full[:, m:m + n].copy(), ddt=full[:, m + n:m + 2 * n].copy(),
                status=full[:, m + 2 * n].astype(np.int32), fit_status=full[:, m + 2 * n + 1].astype(np.int32), cov=cov,
                sqrtW=sqrtW, w_status=w_status)
