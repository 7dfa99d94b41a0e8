"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference in this container.

TEST INFRASTRUCTURE ONLY.  Run as ``python oracle/make_golden.py [names...]`` from
the repo root (needs ``/root/reference``; cannot run on the GPU box).

Each fixture holds the inputs of one reference experiment configuration (first
line of the corresponding ``experiments.sh``), produced by the reference's own
full-order models, noise models and seeds, and the outputs of the reference's
``gpkernels.GP_RBFW`` (``fit`` / ``predict`` / ``compute_lstsq_matrices`` and
``gpr.log_marginal_likelihood``) on them, plus library versions.

What is restated here rather than imported (the reference's ``config*.py`` /
``main.py`` need ``opinf``, which is not installed):
  * the step1 sampling sequence (``ODEs/step1_generate_data.py:69-137``,
    ``PDEs/step1_generate_data.py:15-70``, ``PDEsMulti/step1_generate_data.py:73-123``),
  * configuration constants (``ODEs/config_seird.py:14-17``, ``ODEs/config.py:21-24,92``;
    ``PDEs/config_euler.py:32-39,100-103``, ``PDEs/config.py:89``;
    ``PDEsMulti/config_heat.py:33-49,117-120``, ``PDEsMulti/config.py:84``),
  * POD compression (``PDEs/config_euler.py:50-84``: shift by the mean snapshot,
    divide the three variables by (100, 1e5, 0.1), SVD, project;
    ``PDEsMulti/config_heat.py:69-90``: stack (q, q^2), shift, SVD, project).
"""

from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")


def versions():
    import scipy
    import sklearn

    return dict(numpy=np.__version__, scipy=scipy.__version__, sklearn=sklearn.__version__)


# ------------------------------------------------------------------ data
def data_seird(num_samples=90, noiselevel=0.10, num_regression_points=360, tmax=90):
    """`python3 main.py 090 090 .10 360` (ODEs/experiments.sh:11)."""
    odes, _ = ref_import.load_reference_models()
    np.random.seed(21092023)  # ODEs/config.py:92
    true_parameters = np.array([1.0, 0.25, 0.1, 0.1, 0.05, 0.05])  # config_seird.py:15
    X0 = np.array([0.994, 0.005, 0.001, 0, 0])  # config_seird.py:16
    model = odes.SEIRD2(odes.SEIRD2.convert_parameters(true_parameters))
    model.solve(X0, np.linspace(0, 200, 500))  # true states (keeps solver call order)

    def sample_times():  # ODEs/step1_generate_data.py:69-91 (integersonly=True)
        t = np.random.choice(int(tmax), size=num_samples, replace=False)
        times = np.sort(t).astype(float)
        times[0] = 0
        times[-1] = tmax
        return times

    T, Y = [], []
    for i in range(5):  # synced=False: ODEs/step1_generate_data.py:127-135
        t = sample_times()
        noised = model.noise(model.solve(X0, t), noiselevel)
        T.append(t)
        Y.append(noised[i, :])
    bounds = np.array([(1e-8, 1e5), (0.1, 100), (1e-16, 0.5)])  # ODEs/config.py:21-23
    t_est = np.linspace(0, tmax, num_regression_points)
    return dict(T=np.array(T), Y=np.array(Y), t_est=t_est, bounds=bounds, eta=5e-8,
                seed=21092023, shared_t=False)


def data_euler(num_samples=200, noiselevel=0.03, num_regression_points=400, r=6, tmax=0.06):
    """`python3 main.py 0.06 200 .03 0400 6` (PDEs/experiments.sh:13)."""
    _, pdes = ref_import.load_reference_models()
    np.random.seed(27092023)  # PDEs/config.py:89
    x = np.linspace(0, 2, 201)[:-1]  # config_euler.py:32
    model = pdes.Euler(x)
    q0 = model.initial_conditions(init_params=[22, 20, 24, 95, 105, 100], plot=False)
    model.solve(q0, np.linspace(0, 0.15, 401))  # true states
    t = np.sort(np.random.uniform(0, tmax, size=num_samples))  # step1:48-56
    t[0], t[-1] = 0, tmax
    snaps = model.noise(model.solve(q0, t), noiselevel)
    # POD (config_euler.py:50-84)
    shift = snaps.mean(axis=1)
    Q = snaps - shift[:, None]
    scal = np.array([100.0, 10 * 100.0**2, 1 / 10.0])
    Q = np.concatenate([v / s for v, s in zip(np.split(Q, 3), scal)])
    U, s, _ = np.linalg.svd(Q, full_matrices=False)
    Yc = U[:, :r].T @ Q
    bounds = np.array([(1e-5, 1e5), (1e-5, 1e2), (1e-16, 1e2)])  # config_euler.py:100-102
    t_est = np.linspace(0, tmax, num_regression_points)
    return dict(T=np.tile(t, (r, 1)), Y=Yc, t_est=t_est, bounds=bounds, eta=1e-8,
                seed=27092023, shared_t=True, svdvals=s[:12])


def data_heat(num_samples=20, noiselevel=0.05, num_regression_points=80, r=5, tmax=1.0):
    """`python3 main.py 1 20 .05 80 5` (PDEsMulti/experiments.sh:6)."""
    _, pdes = ref_import.load_reference_models()
    np.random.seed(29012024)  # PDEsMulti/config.py:84
    x = np.linspace(0, 1, 500)
    q0 = pdes.HeatBimodal.initial_conditions(x, 0, 1)
    params = ((-2, 0), (-1, -2), (0, 1), (1, -1), (2, 2))  # config_heat.py:43-49
    tfull = np.linspace(0, 2, 500)
    ts, snaps = [], []
    for (a, b) in params:  # PDEsMulti/step1_generate_data.py:104-121 (synced=False)
        model = pdes.CubicHeatBimodal(x, 0, 1, diffusion=1e-2, a=a, b=b)
        model.solve(q0, tfull)
        t = np.sort(np.random.uniform(0, tmax, size=num_samples))
        t[0], t[-1] = 0, tmax
        snaps.append(model.noise(model.solve(q0, t), noiselevel))
        ts.append(t)
    # POD of (q, q^2) over all trajectories (config_heat.py:69-90; main.py:85-92)
    big = np.hstack(snaps)
    big = np.concatenate((big, big**2))
    shift = big.mean(axis=1)
    U, s, _ = np.linalg.svd(big - shift[:, None], full_matrices=False)
    T, Y = [], []
    for t, Q in zip(ts, snaps):
        Qc = U[:, :r].T @ (np.concatenate((Q, Q**2)) - shift[:, None])
        for i in range(r):
            T.append(t)
            Y.append(Qc[i])
    bounds = np.array([(1e-5, 1e5), (1e-5, 1e2), (1e-16, 1e2)])  # config_heat.py:117-119
    t_est = np.linspace(0, tmax, num_regression_points)
    return dict(T=np.array(T), Y=np.array(Y), t_est=t_est, bounds=bounds, eta=1e-8,
                seed=29012024, shared_t=False, svdvals=s[:12])


# ------------------------------------------------------------------ reference runs
def eval_points(theta_opt, bounds, starts):
    """Fixed-theta evaluation points: optimum, two perturbations, box centre, first restart."""
    lb = np.log(bounds)
    mid = 0.5 * (lb[:, 0] + lb[:, 1])
    pts = [theta_opt, theta_opt + np.array([0.3, -0.2, 0.5]), theta_opt + np.array([-0.5, 0.4, -1.0]),
           0.5 * (theta_opt + mid)]
    pts = [np.clip(p, lb[:, 0], lb[:, 1]) for p in pts]
    if len(starts):
        pts.append(starts[0])
    return np.array(pts)


def run_reference(data, n_restarts, n_cov=2, label=""):
    gpk = ref_import.load_reference_gpkernels()
    T, Y, t_est, bounds, eta = data["T"], data["Y"], data["t_est"], data["bounds"], data["eta"]
    G, m = Y.shape
    lb = np.log(bounds)
    out = dict(starts=[], theta_opt=[], lml_opt=[], alpha_opt=[], pred_mean=[], pred_std=[],
               state_estimate=[], ddt_estimate=[], ddt_covariance=[], thetas_eval=[], lml_eval=[],
               grad_eval=[], cond_eval=[], fit_seconds=[])
    for g in range(G):
        t0 = time.time()
        state = np.random.get_state()
        gp = gpk.GP_RBFW(tuple(bounds[0]), tuple(bounds[1]), tuple(bounds[2]), n_restarts)
        gp.fit(T[g], Y[g])
        after = np.random.get_state()
        # RNG contract: fit() consumes exactly n_restarts draws of 3 uniforms.
        np.random.set_state(state)
        starts = np.array([np.random.uniform(lb[:, 0], lb[:, 1]) for _ in range(n_restarts)]).reshape(-1, 3)
        chk = np.random.get_state()
        assert chk[2] == after[2] and np.array_equal(chk[1], after[1]), "RNG contract violated"
        np.random.set_state(after)
        fit_s = time.time() - t0
        theta = np.array(gp.gpr.kernel_.theta)
        gp.compute_lstsq_matrices(t_est, eta=eta)
        mean, std = gp.predict(t_est)
        pts = eval_points(theta, bounds, starts)
        lmls, grads, conds = [], [], []
        for p in pts:
            l, gr = gp.gpr.log_marginal_likelihood(p, eval_gradient=True)
            K = gp.gpr.kernel_.clone_with_theta(p)(T[g][:, None])
            lmls.append(l)
            grads.append(gr)
            conds.append(np.linalg.cond(K))
        gp.gpr.kernel_.theta = theta
        out["starts"].append(starts)
        out["theta_opt"].append(theta)
        out["lml_opt"].append(gp.gpr.log_marginal_likelihood_value_)
        out["alpha_opt"].append(gp.gpr.alpha_)
        out["pred_mean"].append(mean)
        out["pred_std"].append(std)
        out["state_estimate"].append(gp.state_estimate)
        out["ddt_estimate"].append(gp.ddt_estimate)
        if g < n_cov:
            out["ddt_covariance"].append(gp.ddt_covariance)
        out["thetas_eval"].append(pts)
        out["lml_eval"].append(lmls)
        out["grad_eval"].append(grads)
        out["cond_eval"].append(conds)
        out["fit_seconds"].append(fit_s)
        print(f"[{label}] gp {g}: fit {fit_s:.1f}s theta={theta} lml={out['lml_opt'][-1]:.9f}", flush=True)
    return {k: np.array(v) for k, v in out.items()}


def make_config(name, data_fn, n_restarts=100, n_cov=2):
    data = data_fn()
    res = run_reference(data, n_restarts, n_cov=n_cov, label=name)
    v = versions()
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, f"{name}.npz"),
        T=data["T"], Y=data["Y"], t_est=data["t_est"], bounds=data["bounds"], eta=data["eta"],
        n_restarts=n_restarts, seed=data["seed"], shared_t=data["shared_t"],
        versions=np.array([f"{k}={x}" for k, x in v.items()]), **res,
    )


def make_fixed_theta(name="fixed_theta_synth", sizes=(64, 200, 333, 512, 1024)):
    """LML + gradient of the reference's sklearn regressor at fixed theta, synthetic data
    (SURVEY.md §8d proposal), for sizes that exercise the tiled large-matrix path."""
    gpk = ref_import.load_reference_gpkernels()
    from gp_oracle import synthetic_trajectories

    bounds = np.array([(1e-5, 1e5), (1e-5, 1e2), (1e-16, 1e2)])
    thetas = np.log(np.array([
        [1.0, 0.1, 1e-3],
        [2.5, 0.05, 1e-2],
        [0.7, 0.3, 3e-3],
        [10.0, 0.02, 1e-1],
        [1.0, 0.1, 1e-14],   # not positive definite in FP64 at these sizes -> status parity
    ]))
    save = dict(thetas=thetas, sizes=np.array(sizes))
    for m in sizes:
        t, y = synthetic_trajectories(2, m, seed=m)
        lml = np.zeros((2, len(thetas)))
        grad = np.zeros((2, len(thetas), 3))
        cond = np.zeros((2, len(thetas)))
        for g in range(2):
            gp = gpk.GP_RBFW(tuple(bounds[0]), tuple(bounds[1]), tuple(bounds[2]), 0)
            gp.gpr.optimizer = None  # fixed-theta use only: skip the optimiser
            gp.fit(t, y[g])
            for k, th in enumerate(thetas):
                l, gr = gp.gpr.log_marginal_likelihood(th, eval_gradient=True)
                lml[g, k], grad[g, k] = l, gr
                K = gp.gpr.kernel_.clone_with_theta(th)(t[:, None])
                cond[g, k] = np.linalg.cond(K)
            print(f"[fixed] m={m} g={g} lml={lml[g]}", flush=True)
        save[f"t_{m}"], save[f"y_{m}"] = t, y
        save[f"lml_{m}"], save[f"grad_{m}"], save[f"cond_{m}"] = lml, grad, cond
    v = versions()
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"{name}.npz"),
                        versions=np.array([f"{k}={x}" for k, x in v.items()]), **save)


def _extreme_cond(K):
    """cond_2 of a symmetric positive definite K from its extreme eigenvalues."""
    import scipy.linalg as sla

    m = K.shape[0]
    lo = sla.eigvalsh(K, subset_by_index=[0, 0], check_finite=False)[0]
    hi = sla.eigvalsh(K, subset_by_index=[m - 1, m - 1], check_finite=False)[0]
    return float(hi / lo) if lo > 0 else float("inf")


LARGE_THETAS = np.log(np.array([
    [1.0, 0.1, 1e-3],
    [2.5, 0.05, 1e-2],
    [0.7, 0.3, 3e-3],
]))


def make_fixed_theta_large(name="fixed_theta_large", ref_sizes=(4096, 8192), lean_sizes=(16384,), truth=True,
                           moments_m=4096, moments_n=512):
    """The benchmarked sizes (BASELINE configs[3], [4] and the north star's n = 8192), GP 0 of the synthetic workload.

    * m in ref_sizes: LML + gradient from the UNMODIFIED reference (GP_RBFW.gpr.log_marginal_likelihood);
    * m in lean_sizes: the reference's m x m x 3 gradient tensor (6.4 GB at 16384, ~5 copies live) does not fit a
      test-generation budget, so ``gp_oracle.np_lml_grad_lean`` (same LAPACK calls, traces by row blocks) is used;
      it is checked against the unmodified reference at m = 4096 here and the agreement is stored;
    * every (m, theta): the 80-bit long-double value of ``oracle/lml_ld.c`` ("truth") and the reference's own error
      against it, so that a GPU test can ask "is the CUDA result as close to the truth as LAPACK's?";
    * posterior moments (predict, compute_lstsq_matrices) of the unmodified reference at m = moments_m.
    Only theta, LML, grad, alpha, cond and the inputs are stored."""
    gpk = ref_import.load_reference_gpkernels()
    import gp_oracle as orc

    bounds = np.array([(1e-5, 1e5), (1e-5, 1e2), (1e-16, 1e2)])
    thetas = LARGE_THETAS
    save = dict(thetas=thetas, sizes=np.array(list(ref_sizes) + list(lean_sizes)),
                ref_sizes=np.array(ref_sizes), lean_sizes=np.array(lean_sizes))

    def new_gp(t, y, th):
        gp = gpk.GP_RBFW(tuple(bounds[0]), tuple(bounds[1]), tuple(bounds[2]), 0)
        gp.gpr.optimizer = None          # fixed-theta use only
        gp.gpr.kernel.theta = th          # fit() then factors K(th): alpha_, L_ at th
        gp.fit(t, y)
        return gp

    for m in list(ref_sizes) + list(lean_sizes):
        t, y = orc.synthetic_trajectories(2, m, seed=m)
        y0 = y[0]
        nth = len(thetas) if m in ref_sizes else 2
        lml, grad, cond = np.zeros(nth), np.zeros((nth, 3)), np.zeros(nth)
        alpha = np.zeros((nth, m))
        lean_gap = np.zeros((nth, 2))
        for k in range(nth):
            th = thetas[k]
            t0 = time.time()
            if m in ref_sizes:
                gp = new_gp(t, y0, th)
                l, gr = gp.gpr.log_marginal_likelihood(th, eval_gradient=True)
                alpha[k] = gp.gpr.alpha_
                K = gp.gpr.kernel_(t[:, None])
                cond[k] = _extreme_cond(K)
                del K
                if m == min(ref_sizes):
                    l2, g2, _ = orc.np_lml_grad_lean(t, y0, th)
                    lean_gap[k] = abs(l2 - l) / abs(l), np.abs(g2 - gr).max() / np.abs(gr).max()
                del gp
            else:
                l, gr, st, al = orc.np_lml_grad_lean(t, y0, th, want_alpha=True)
                assert st == 0
                alpha[k] = al
                cond[k] = _extreme_cond(orc.np_kernel(t, th))
            lml[k], grad[k] = l, gr
            print(f"[large] m={m} k={k} lml={l!r} grad={gr} cond={cond[k]:.3e} ({time.time()-t0:.1f}s)", flush=True)
        save[f"t_{m}"], save[f"y_{m}"] = t, y0
        save[f"lml_{m}"], save[f"grad_{m}"], save[f"cond_{m}"], save[f"alpha_{m}"] = lml, grad, cond, alpha
        if m == min(ref_sizes):
            save["lean_vs_reference_rel"] = lean_gap

    # posterior moments of the unmodified reference at the benchmarked size
    m = moments_m
    t, y = orc.synthetic_trajectories(2, m, seed=m)
    th = thetas[0]
    gp = new_gp(t, y[0], th)
    t_est = np.linspace(0.0, 1.0, moments_n)
    mean, std = gp.predict(t_est)
    gp.compute_lstsq_matrices(t_est, eta=1e-8)
    save.update(mom_m=m, mom_theta=th, mom_t_est=t_est, mom_pred_mean=mean, mom_pred_std=std,
                mom_state=gp.state_estimate, mom_ddt=gp.ddt_estimate, mom_cov_diag=np.diag(gp.ddt_covariance).copy(),
                mom_cov_sub=gp.ddt_covariance[::8, ::8].copy())
    print("[large] moments done", flush=True)
    v = versions()
    path = os.path.join(GOLDEN_DIR, f"{name}.npz")
    np.savez_compressed(path, versions=np.array([f"{k}={x}" for k, x in v.items()]), **save)

    if truth:
        for m in list(ref_sizes) + list(lean_sizes):
            nth = len(save[f"lml_{m}"])
            tl, tg, ta = np.zeros(nth), np.zeros((nth, 3)), np.zeros((nth, m))
            for k in range(nth):
                t0 = time.time()
                tl[k], tg[k], ta[k] = orc.ld_truth(save[f"t_{m}"], save[f"y_{m}"], thetas[k])
                print(f"[truth] m={m} k={k} lml={tl[k]!r} ref_err={abs(tl[k]-save[f'lml_{m}'][k])/abs(tl[k]):.2e} "
                      f"grad_err={np.abs(tg[k]-save[f'grad_{m}'][k]).max()/np.abs(tg[k]).max():.2e} "
                      f"alpha_err={np.abs(ta[k]-save[f'alpha_{m}'][k]).max()/np.abs(ta[k]).max():.2e} "
                      f"({time.time()-t0:.1f}s)", flush=True)
                save[f"truth_lml_{m}"], save[f"truth_grad_{m}"], save[f"truth_alpha_{m}"] = tl, tg, ta
                np.savez_compressed(path, versions=np.array([f"{k2}={x}" for k2, x in v.items()]), **save)


ALL = dict(
    fixed=lambda: make_fixed_theta(),
    large=lambda: make_fixed_theta_large(),
    seird=lambda: make_config("seird_090_090_10_360", data_seird),
    heat=lambda: make_config("heat_1_20_05_80_5", data_heat),
    euler=lambda: make_config("euler_006_200_03_400_6", data_euler),
    # sparse-data line of ODEs/experiments.sh (`main.py 120 010 .05 480`): only 10 samples per state
    seird_sparse=lambda: make_config("seird_120_010_05_480", lambda: data_seird(10, 0.05, 480, 120)),
    # sparse-data line of PDEs/experiments.sh (`main.py 0.06 50 .01 0400 6`): 50 samples, 1 % noise
    euler_sparse=lambda: make_config("euler_006_050_01_400_6", lambda: data_euler(50, 0.01, 400, 6, 0.06), n_cov=1),
)

if __name__ == "__main__":
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    os.chdir("/tmp")
    names = sys.argv[1:] or [n for n in ALL if n != "large"]   # 'large' takes ~1 h: ask for it by name
    for n in names:
        t0 = time.time()
        ALL[n]()
        print(f"== {n} done in {time.time()-t0:.1f}s", flush=True)
