"""world_size-2 gloo test of the (GP x start) sharding + all-gather logic, with a CPU stand-in engine."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


class FakeEngine:
    """Deterministic stand-in for _lib.Context: results are pure functions of the inputs, so the sharded
    run must reproduce the single-process run exactly.  The "log-marginal likelihood" is a smooth non-quadratic
    function with a different optimum per GP, so the L-BFGS-B runs have different lengths (pairs drop out of the
    lock-step rounds at different times, which is what the per-round re-balancing has to cope with); one GP's
    objective is -inf away from its start region (the non-PD case)."""

    def __init__(self):
        self.evaluated = 0

    def upload_problem(self, T, Y):
        self.T, self.Y = np.array(T), np.array(Y)

    def lml_grad_resident(self, theta, gp_of):
        theta = np.atleast_2d(theta)
        self.evaluated += theta.shape[0]
        c = self.Y[gp_of, :3] * 0.5
        w = 1.0 + np.arange(3)[None, :] * (1 + np.asarray(gp_of)[:, None])
        d = theta - c
        lml = -(w * d ** 2).sum(1) - 0.1 * (d ** 4).sum(1) + self.T[gp_of].sum(1)
        grad = -2 * w * d - 0.4 * d ** 3
        bad = (np.asarray(gp_of) == 3) & (theta[:, 2] < -4.0)
        lml = np.where(bad, -np.inf, lml)
        grad = np.where(bad[:, None], 0.0, grad)
        return lml, grad, bad.astype(np.int32)

    def fit(self, T, Y, bounds_log, starts, gp_of=None, opts=None):
        """Single-process driver: the same optimiser pool, all live pairs evaluated here every round."""
        from gpbo_pkg import pkg

        self.upload_problem(T, Y)
        pool = pkg._lib.OptimizerPool(bounds_log, starts, opts)
        while True:
            idx, theta = pool.live()
            if idx.size == 0:
                break
            lml, grad, _ = self.lml_grad_resident(theta, np.asarray(gp_of)[idx])
            pool.feed(idx, lml, grad)
        return pool.result()

    def lstsq_moments(self, T, Y, theta, t_est, want_cov=True):
        G = T.shape[0]
        state = theta[:, :1] * t_est[None, :]
        ddt = Y.sum(1)[:, None] + t_est[None, :]
        cov = np.einsum("g,i,j->gij", theta[:, 1], t_est, t_est) if want_cov else None
        return state, ddt, cov, np.zeros(G, np.int32)

    def lstsq_weights(self, T, Y, theta, t_est, eta):
        state, ddt, cov, st = self.lstsq_moments(T, Y, theta, t_est)
        n = t_est.shape[-1]
        w = cov + eta * np.eye(n)[None]
        G = T.shape[0]
        return state, ddt, cov, w, st, (np.arange(G) % 2).astype(np.int32), np.full(G, 9, np.int32)

    def predict(self, T, Y, theta, t_star, want_alpha=False):
        G = T.shape[0]
        alpha = Y * theta[:, 2:3]
        return np.zeros((G, 1)), np.zeros((G, 1)), alpha, np.zeros(G, np.int32)


def _problem():
    rng = np.random.default_rng(5)
    G, m, S = 5, 12, 7
    T = np.sort(rng.uniform(0, 1, (G, m)), axis=1)
    Y = rng.standard_normal((G, m))
    bl = np.array([[-3.0, 3.0], [-2.0, 2.0], [-5.0, 1.0]])
    starts = rng.uniform(bl[:, 0], bl[:, 1], size=(G * S, 3))
    gp_of = np.repeat(np.arange(G, dtype=np.int32), S)
    return T, Y, bl, starts, gp_of


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from gpbo_pkg import pkg

    dist.init_process_group("gloo", rank=rank, world_size=world)
    T, Y, bl, starts, gp_of = _problem()
    eng = FakeEngine()
    live_per_round = []
    res = pkg.sharding.fit_pairs(eng, T, Y, bl, starts, gp_of, group=True, stats=live_per_round)
    # every round this rank evaluated its block of the live pairs: the work is balanced to within one pair per round
    b = [pkg.sharding.block_bounds(n, world) for n in live_per_round]
    assert eng.evaluated == sum(int(x[rank + 1] - x[rank]) for x in b) == res["local_evals"]
    assert res["evals"] == sum(live_per_round) and res["rounds"] == len(live_per_round)
    assert live_per_round[0] == starts.shape[0] and live_per_round[-1] < live_per_round[0]
    G = T.shape[0]
    funs = res["fun"].reshape(G, -1)
    theta_opt = res["theta"].reshape(G, -1, 3)[np.arange(G), funs.argmin(1)]
    mom = pkg.sharding.moments(eng, T, Y, theta_opt, np.linspace(0, 1, 6), group=True)
    owned = [g for g in range(G) if mom["cov"][g] is not None]
    # with eta: covariance AND sqrtW stay with the owner, computed by one lstsq_weights call per rank
    momw = pkg.sharding.moments(eng, T, Y, theta_opt, np.linspace(0, 1, 6), group=True, eta=0.5)
    owned_w = [g for g in range(G) if momw["sqrtW"][g] is not None]
    assert owned_w == owned and all(np.array_equal(momw["sqrtW"][g], momw["cov"][g] + 0.5 * np.eye(6)) for g in owned)
    assert np.array_equal(momw["state"], mom["state"]) and np.array_equal(momw["alpha"], mom["alpha"])
    # gather_cov: every rank gets every covariance / sqrtW (what the reference's step 3 reads on one process)
    momg = pkg.sharding.moments(eng, T, Y, theta_opt, np.linspace(0, 1, 6), group=True, eta=0.5, gather_cov=True)
    assert all(momg["cov"][g] is not None and momg["sqrtW"][g] is not None for g in range(G))
    for g in owned:
        assert np.array_equal(momg["cov"][g], momw["cov"][g]) and np.array_equal(momg["sqrtW"][g], momw["sqrtW"][g])
    # statuses are all-gathered (the fake engine reports local index % 2): the same vector on every rank
    assert np.array_equal(momg["w_status"], (np.arange(G) // world) % 2)
    q.put((rank, res["theta"], res["fun"], res["status"], mom["alpha"], mom["state"], mom["ddt"], owned,
           res["nfev"], np.array([momg["cov"][g] for g in range(G)])))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_fit_and_moments_match_single_process():
    from gpbo_pkg import pkg

    T, Y, bl, starts, gp_of = _problem()
    eng = FakeEngine()
    ref = pkg.sharding.fit_pairs(eng, T, Y, bl, starts, gp_of, group=None)
    G = T.shape[0]
    theta_opt = ref["theta"].reshape(G, -1, 3)[np.arange(G), ref["fun"].reshape(G, -1).argmin(1)]
    refm = pkg.sharding.moments(eng, T, Y, theta_opt, np.linspace(0, 1, 6), group=None)

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owned_all = []
    assert len(set(ref["nfev"].tolist())) > 2            # the runs really have different lengths
    assert (ref["status"] == 5).sum() >= 0
    covs = [o[9] for o in outs]
    assert np.array_equal(covs[0], covs[1])
    for rank, theta, fun, status, alpha, state, ddt, owned, nfev, _ in outs:
        assert np.array_equal(theta, ref["theta"]) and np.array_equal(fun, ref["fun"])
        assert np.array_equal(status, ref["status"]) and np.array_equal(nfev, ref["nfev"])
        assert np.array_equal(alpha, refm["alpha"]) and np.array_equal(state, refm["state"])
        assert np.array_equal(ddt, refm["ddt"])
        assert owned == list(range(rank, G, 2))      # covariance stays on the owning rank
        owned_all += owned
    assert sorted(owned_all) == list(range(G))


def test_shard_indices_cover_everything_once():
    from gpbo_pkg import pkg

    for n in (0, 1, 7, 64, 2112):
        for w in (1, 2, 4, 8):
            allidx = np.concatenate([pkg.sharding.shard_indices(n, r, w) for r in range(w)])
            assert sorted(allidx.tolist()) == list(range(n))
