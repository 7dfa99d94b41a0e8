"""Multi-GPU sharding of the step2_fitgps batch.

The reference is single-process; here the (GP x start) pairs are independent optimisation problems and the GPs are
independent for prediction, so one process per GPU works on a slice with no data-path collective.

* **Optimisation** (``fit_pairs``).  L-BFGS-B runs have heavy-tailed lengths (one start in a thousand needs 100x the
  average number of evaluations), so a static split of the pairs leaves ranks idle.  Instead every rank keeps the
  whole optimiser pool (``_lib.OptimizerPool``: B state machines of ~800 B, pure host code, deterministic) and the
  *evaluations* are re-balanced every lock-step round: the pairs still running are cut into ``world`` contiguous,
  equally sized slices, each rank evaluates LML + gradient of its slice on its GPU against the problem resident in
  its HBM, one all-gather moves ``[lml, grad(3)]`` per live pair (32 B), and every rank feeds every optimiser.  All
  ranks therefore hold identical optimiser states, nothing ever migrates, and a round costs
  ``ceil(live / world)`` evaluations on every rank.
* **Selection**: per-GP argmin of sklearn ``_gpr.py:336-340`` -- every rank has every result.
* **Moments** (``moments``): GPs ``rank::world`` per rank, one all-gather of ``alpha_, state_estimate, ddt_estimate``;
  ``ddt_covariance`` and ``sqrtW`` (m'^2 per GP each) stay on the owning rank unless ``gather_cov`` asks for them on
  every rank (what the reference's step 3 reads: ``PDEs/step3_estimate.py:212``, ``PDEs/main.py:219``).

The collectives go through ``torch.distributed`` (NCCL over NVLink on GPUs; gloo in the CPU tests).  ``engine`` is
anything with the ``_lib.Context`` methods ``fit`` / ``upload_problem`` / ``lml_grad_resident`` / ``lstsq_moments`` /
``lstsq_weights`` / ``predict``.
"""

from __future__ import annotations

import numpy as np

from . import _lib


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


def block_bounds(n, world):
    """Contiguous, equally sized slices of n items: rank r owns [b[r], b[r+1])."""
    return (np.arange(world + 1) * n) // world


def _device_for(dist, group):
    import torch

    g = None if group is True else group
    if dist.get_backend(g) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def _all_gather_rows(dist, group, local, n_total, rank, world):
    """All-gather row blocks of a cyclic partition back into global order.  local: (n_local, k) float64."""
    import torch

    g = None if group is True else group
    k = local.shape[1]
    per = (n_total + world - 1) // world
    dev = _device_for(dist, group)
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


class _RoundGather:
    """All-gather of the per-round evaluations over a block partition, with buffers reused across rounds."""

    def __init__(self, dist, group, world, n_max, k):
        import torch

        self.dist, self.g, self.world, self.k = dist, (None if group is True else group), world, k
        self.dev = _device_for(dist, group)
        per = (n_max + world - 1) // world
        self.host = torch.zeros((per, k), dtype=torch.float64)
        if self.dev.type == "cuda":
            self.host = self.host.pin_memory()
        self.send = torch.zeros((per, k), dtype=torch.float64, device=self.dev)
        self.recv = torch.zeros((world * per, k), dtype=torch.float64, device=self.dev)
        self.per_max = per

    def __call__(self, local, n):
        """local: (n_local, k) rows of this rank's slice of the n live pairs -> (n, k) rows of all ranks."""
        b = block_bounds(n, self.world)
        per = int((b[1:] - b[:-1]).max()) if n else 0
        if per == 0:
            return np.zeros((0, self.k))
        self.host[: local.shape[0]] = torch_from(local)
        send = self.send[:per]
        send.copy_(self.host[:per], non_blocking=True)
        recv = self.recv[: self.world * per]
        self.dist.all_gather_into_tensor(recv, send, group=self.g)
        allr = recv.cpu().numpy().reshape(self.world, per, self.k)
        return np.concatenate([allr[r, : b[r + 1] - b[r]] for r in range(self.world)], axis=0)


def torch_from(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))


def fit_pairs(engine, T, Y, bounds_log, starts, gp_of, group=None, opts=None, pool_factory=None, stats=None,
              max_rounds=None):
    """Optimise all (GP, start) pairs.  Single process: ``engine.fit`` (the library's own lock-step / persistent
    driver).  With a process group: replicated optimiser pool, evaluations re-balanced over the ranks every round
    (module docstring).  Returns the dict of ``_lib.Context.fit`` -- identical on every rank; ``evals`` counts the
    evaluations of all ranks, ``local_evals`` this rank's share.  ``max_rounds`` stops after that many lock-step
    rounds (benchmarks); pairs still running then report their last accepted iterate with status -1."""
    dist, rank, world = _dist(group)
    if world == 1 and max_rounds is None:
        return engine.fit(T, Y, bounds_log, starts, gp_of=gp_of, opts=opts)
    gp_of = np.ascontiguousarray(gp_of, dtype=np.int32)
    if world > 1 and max_rounds is None and T.shape[1] <= getattr(engine, "small_max", 0):
        # reference-size problems: the library runs a whole fit as one persistent kernel with the optimisers on the
        # device (no rounds to re-balance; a straggler only occupies its own CTA) -> static cyclic split of the pairs
        B = starts.shape[0]
        idx = shard_indices(B, rank, world)
        if idx.size:
            loc = engine.fit(T, Y, bounds_log, starts[idx], gp_of=gp_of[idx], opts=opts)
            packed = np.column_stack([loc["theta"], loc["fun"], loc["nfev"], loc["nit"], loc["status"]])
            local_evals = int(loc["evals"])
        else:
            packed, local_evals = np.zeros((0, 7)), 0
        full = _all_gather_rows(dist, group, packed, B, rank, world)
        nfev = full[:, 4].astype(np.int32)
        return dict(theta=full[:, :3].copy(), fun=full[:, 3].copy(), nfev=nfev, nit=full[:, 5].astype(np.int32),
                    status=full[:, 6].astype(np.int32), evals=int(nfev.sum()), rounds=int(nfev.max()) if B else 0,
                    local_evals=local_evals)
    pool = (pool_factory or _lib.OptimizerPool)(bounds_log, starts, opts)
    engine.upload_problem(T, Y)
    gather = _RoundGather(dist, group, world, starts.shape[0], 4) if world > 1 else None
    local_evals = 0
    rounds = 0
    while max_rounds is None or rounds < max_rounds:
        idx, theta = pool.live()
        n = idx.size
        if n == 0:
            break
        rounds += 1
        b = block_bounds(n, world)
        lo, hi = int(b[rank]), int(b[rank + 1])
        lml, grad, _ = engine.lml_grad_resident(theta[lo:hi], gp_of[idx[lo:hi]])
        local_evals += hi - lo
        loc = np.column_stack([lml, grad]) if hi > lo else np.zeros((0, 4))
        full = gather(loc, n) if gather is not None else loc
        pool.feed(idx, full[:, 0], full[:, 1:4])
        if stats is not None:
            stats.append(n)
    res = pool.result()
    res["local_evals"] = local_evals
    if hasattr(pool, "close"):
        pool.close()
    return res


def moments(engine, T, Y, theta_opt, t_est, group=None, want_cov=True, eta=None, gather_cov=False, chunk=None,
            keep_cov=True):
    """alpha_, state/ddt estimates, derivative covariance (and, when ``eta`` is given, sqrtW) of every GP;
    GPs ``rank::world`` per rank.  ``cov[g]`` / ``sqrtW[g]`` are None on ranks that do not own GP g unless
    ``gather_cov`` (then every rank receives them: G * n^2 doubles each over the interconnect).  The numerical
    statuses (``status``, ``fit_status``, ``w_status``) are all-gathered, so every rank raises the same errors.
    The owned GPs are processed ``chunk`` at a time (default: about 4 GB of covariance per call) to bound the
    staging memory on device and host; ``keep_cov=False`` drops each chunk's matrices after the call (benchmarks)."""
    dist, rank, world = _dist(group)
    G, m = T.shape
    n = t_est.shape[-1]
    idx = shard_indices(G, rank, world)
    cov = [None] * G
    sqrtW = [None] * G
    if chunk is None:
        chunk = max(1, int(4e9 // (8.0 * n * n * (2 if eta is not None else 1))))
    if idx.size:
        rows = []
        for c0 in range(0, idx.size, chunk):
            sub = idx[c0:c0 + chunk]
            pts = t_est if t_est.ndim == 1 else t_est[sub]
            if eta is not None:
                state, ddt, c, w, st, wst, _ = engine.lstsq_weights(T[sub], Y[sub], theta_opt[sub], pts, eta)
            else:
                state, ddt, c, st = engine.lstsq_moments(T[sub], Y[sub], theta_opt[sub], pts, want_cov=want_cov)
                w, wst = None, np.zeros(sub.size, dtype=np.int32)
            _, _, alpha, fst = engine.predict(T[sub], Y[sub], theta_opt[sub], T[sub][:, :1], want_alpha=True)
            if keep_cov:
                for k, g in enumerate(sub):
                    cov[g] = c[k] if c is not None else None
                    if w is not None:
                        sqrtW[g] = w[k]
            rows.append(np.column_stack([alpha, state, ddt, st, fst, wst]))
        packed = np.vstack(rows)
    else:
        packed = np.zeros((0, m + 2 * n + 3))
    if world == 1:
        full = packed
    else:
        full = _all_gather_rows(dist, group, packed, G, rank, world)
        if gather_cov:
            wanted = [("cov", cov)] if (want_cov or eta is not None) else []
            if eta is not None:
                wanted.append(("sqrtW", sqrtW))
            for _, store in wanted:
                loc = np.array([store[g].ravel() for g in idx]) if idx.size else np.zeros((0, n * n))
                allm = _all_gather_rows(dist, group, loc, G, rank, world)
                for g in range(G):
                    store[g] = allm[g].reshape(n, n)
    return dict(alpha=full[:, :m].copy(), state=full[:, m:m + n].copy(), ddt=full[:, m + n:m + 2 * n].copy(),
                status=full[:, m + 2 * n].astype(np.int32), fit_status=full[:, m + 2 * n + 1].astype(np.int32),
                w_status=full[:, m + 2 * n + 2].astype(np.int32), cov=cov, sqrtW=sqrtW)
