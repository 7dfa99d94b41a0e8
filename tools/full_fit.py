"""Full BASELINE configs[3] fit on one B200: r = 64 modes x m = 4096 points x 32 starts (start 0 at theta = 0 plus 31
drawn uniformly in the log-box, sklearn n_restarts_optimizer = 31), multi-start L-BFGS-B in lock-step, then posterior
moments (state, ddt, ddt covariance) and sqrtW at m' = m estimation points.  Prints one JSON line."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from gpbo_pkg import pkg

r = int(sys.argv[1]) if len(sys.argv) > 1 else 64
m = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
S = int(sys.argv[3]) if len(sys.argv) > 3 else 32
ctx = pkg.default_context(0)
t, Y = pkg.workload.synthetic_trajectories(r, m, seed=0)
T = np.tile(t, (r, 1))
b = np.log(np.array([(1e-5, 1e5), (1e-5, 1e2), (1e-16, 1e2)]))     # Euler/Heat hyper-parameter box (config_euler.py:100-103)
rng = np.random.default_rng(7)
starts = rng.uniform(b[:, 0], b[:, 1], size=(r * S, 3))
starts[::S] = 0.0
gp_of = np.repeat(np.arange(r, dtype=np.int32), S)
t0 = time.perf_counter()
res = ctx.fit(T, Y, b, starts, gp_of)
t_fit = time.perf_counter() - t0
funs = np.where(np.isfinite(res["fun"]), res["fun"], np.inf).reshape(r, S)
best = res["theta"].reshape(r, S, 3)[np.arange(r), funs.argmin(1)]
t_est = np.linspace(0, 1, m)
t1 = time.perf_counter()
mean, std, alpha, st = ctx.predict(T, Y, best, t_est, want_alpha=True)
t_pred = time.perf_counter() - t1
t2 = time.perf_counter()
for g0 in range(0, r, 16):          # posterior moments WITHOUT sqrtW: part of the fits/s definition
    sl = slice(g0, min(r, g0 + 16))
    ctx.lstsq_moments(T[sl], Y[sl], best[sl], t_est)
t_lstsq = time.perf_counter() - t2
t3 = time.perf_counter()
wst_all, wit_all = [], []
for g0 in range(0, r, 16):          # 16 GPs per call keeps the host copies of cov + sqrtW at 4.3 GB
    sl = slice(g0, min(r, g0 + 16))
    state, ddt, cov, w, st2, wst, wit = ctx.lstsq_weights(T[sl], Y[sl], best[sl], t_est, 1e-8)
    wst_all += list(wst)
    wit_all += list(wit)
t_mom = time.perf_counter() - t3
nfev = res["nfev"]
print(json.dumps({
    "workload": f"r={r} modes x m={m} x {S} starts, m'={m}", "fit_seconds": t_fit, "predict_seconds": t_pred,
    "lstsq_moments_seconds": t_lstsq, "lstsq_weights_seconds": t_mom,
    "fits_per_s": r / (t_fit + t_pred + t_lstsq),            # all starts + all posterior moments, sqrtW excluded
    "fits_per_s_with_sqrtw": r / (t_fit + t_pred + t_mom), "lml_grad_evals": res["evals"], "rounds": res["rounds"],
    "evals_per_s_during_fit": res["evals"] / t_fit, "nfev_mean": float(nfev.mean()), "nfev_max": int(nfev.max()),
    "best_lml_min_max": [float((-funs.min(1)).min()), float((-funs.min(1)).max())],
    "opt_status_counts": {int(k): int(v) for k, v in zip(*np.unique(res["status"], return_counts=True))},
    "sqrtw_status_ok": int(sum(1 for x in wst_all if x == 0)), "sqrtw_iters_max": int(max(wit_all))}))
