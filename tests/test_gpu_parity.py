"""Parity of the CUDA path (through the C ABI) with the oracle and the reference's golden vectors.

Tolerances (BASELINE.json north_star): at fixed hyper-parameters LML / gradient / posterior moments within
1e-10 relative in FP64 for well-conditioned kernel matrices (cond(K) <= 1e6; looser, stated per test, above
that, because FP64 LAPACK itself is not 1e-10-accurate there -- SURVEY.md §7 hard part 4); gradients are
compared relative to max(1, |g|_inf) (they cancel to ~1e-7 at an optimum); optimised LML within 1e-8 relative.
"""
import numpy as np
import pytest

import gp_oracle as orc
from conftest import REAL_CONFIGS, load_golden, record

pytestmark = pytest.mark.gpu


def rel(a, b, scale=None):
    a, b = np.asarray(a), np.asarray(b)
    s = np.abs(b).max() if scale is None else scale
    return np.abs(a - b).max() / max(s, 1e-300)


def tol_for(cond):
    # two backward-stable factorizations may differ by a modest multiple of cond * eps
    return 1e-10 if cond <= 1e6 else min(1e-4, 1e-15 * cond)


# ------------------------------------------------------------------ LML + gradient
@pytest.mark.parametrize("m", [64, 200, 333, 512, 1024])
def test_lml_grad_fixed_theta_golden(ctx, m):
    g = load_golden("fixed_theta_synth")
    t, y, thetas = g[f"t_{m}"], g[f"y_{m}"], g["thetas"][:4]
    T = np.tile(t, (2, 1))
    theta = np.tile(thetas, (2, 1))
    gp_of = np.repeat(np.arange(2, dtype=np.int32), len(thetas))
    lml, grad, st = ctx.lml_grad(T, y, theta, gp_of)
    k = 0
    for gi in range(2):
        for j in range(len(thetas)):
            cond = g[f"cond_{m}"][gi, j]
            tol = tol_for(cond)
            ref_l, ref_g = g[f"lml_{m}"][gi, j], g[f"grad_{m}"][gi, j]
            assert st[k] == 0
            if cond <= 1e6:
                record(f"lml_rel[synth m<=1024, cond<=1e6, {ctx.path}]", abs(lml[k] - ref_l) / abs(ref_l))
                record(f"grad_rel[synth m<=1024, cond<=1e6, {ctx.path}]", rel(grad[k], ref_g, max(1.0, np.abs(ref_g).max())))
            assert abs(lml[k] - ref_l) <= tol * abs(ref_l), (m, gi, j, cond, lml[k], ref_l)
            assert rel(grad[k], ref_g, max(1.0, np.abs(ref_g).max())) <= 10 * tol, (m, gi, j, cond, grad[k], ref_g)
            k += 1


@pytest.mark.parametrize("name", REAL_CONFIGS)
def test_lml_grad_real_configs_golden(ctx, name):
    g = load_golden(name)
    T, Y = g["T"], g["Y"]
    G, P = g["thetas_eval"].shape[:2]
    theta = g["thetas_eval"].reshape(-1, 3)
    gp_of = np.repeat(np.arange(G, dtype=np.int32), P)
    lml, grad, st = ctx.lml_grad(T, Y, theta, gp_of)
    checked = 0
    for gi in range(G):
        for j in range(P):
            k = gi * P + j
            cond, ref_l, ref_g = g["cond_eval"][gi, j], g["lml_eval"][gi, j], g["grad_eval"][gi, j]
            if cond > 1e10 or not np.isfinite(ref_l):
                continue
            tol = tol_for(cond)
            assert st[k] == 0
            if cond <= 1e6:
                record(f"lml_rel[real configs, cond<=1e6, {ctx.path}]", abs(lml[k] - ref_l) / max(1.0, abs(ref_l)))
                record(f"grad_rel[real configs, cond<=1e6, {ctx.path}]", rel(grad[k], ref_g, max(1.0, np.abs(ref_g).max())))
            if cond <= 1e6 and "truth_lml_eval" in g.files and np.isfinite(g["truth_lml_eval"][gi, j]):
                # 80-bit truth (oracle/lml_ld.c): the CUDA result must be within 1e-10 of it, or -- where the
                # reference's own FP64 LAPACK path is further off than that -- at most 3x as far as the reference
                tl, tg = g["truth_lml_eval"][gi, j], g["truth_grad_eval"][gi, j]
                # gradient scale: the largest gradient among this GP's evaluation points -- at the fitted optimum
                # (j = 0) the gradient itself is ~1e-6 by cancellation of O(100) traces, and the optimiser's own
                # resolution is pgtol = 1e-5 absolute
                gs = max(1.0, np.nanmax(np.abs(g["truth_grad_eval"][gi])))
                cl, ll = abs(lml[k] - tl) / max(1.0, abs(tl)), abs(ref_l - tl) / max(1.0, abs(tl))
                cg, lg = rel(grad[k], tg, gs), rel(ref_g, tg, gs)
                record(f"lml_vs_truth[real configs, {ctx.path}] cuda", cl)
                record(f"lml_vs_truth[real configs, {ctx.path}] lapack", ll)
                record(f"grad_vs_truth[real configs, {ctx.path}] cuda", cg)
                record(f"grad_vs_truth[real configs, {ctx.path}] lapack", lg)
                assert cl <= max(1e-10, 3 * ll), (gi, j, cond, cl, ll)
                assert cg <= max(1e-10, 3 * lg), (gi, j, cond, cg, lg)
            assert abs(lml[k] - ref_l) <= tol * max(1.0, abs(ref_l)), (gi, j, cond, lml[k], ref_l)
            assert rel(grad[k], ref_g, max(1.0, np.abs(ref_g).max())) <= 10 * tol, (gi, j, cond, grad[k], ref_g)
            checked += 1
    assert checked >= 3 * G


def test_lml_grad_against_oracle_m2048(ctx):
    t, y = orc.synthetic_trajectories(1, 2048, seed=11)
    th = np.log([2.5, 0.05, 1e-2])
    lml, grad, st = ctx.lml_grad(t[None], y, th[None])
    l0, g0, _ = orc.np_lml_grad(t, y[0], th)
    assert st[0] == 0
    assert abs(lml[0] - l0) <= 1e-10 * abs(l0)
    assert rel(grad[0], g0, max(1.0, np.abs(g0).max())) <= 1e-9


def test_not_positive_definite_status(ctx):
    """All-equal abscissae with chi below one ulp of sigma^2: exact zero pivot in LAPACK and here alike
    (sklearn returns (-inf, 0), _gpr.py:589-593)."""
    t = np.full(64, 0.5)
    y = np.linspace(-1, 1, 64)
    th = np.log([1.0, 0.1, 1e-17])
    l0, g0, s0 = orc.np_lml_grad(t, y, th)
    assert s0 == 1 and l0 == -np.inf
    lml, grad, st = ctx.lml_grad(t[None], y[None], th[None])
    assert st[0] == 1 and lml[0] == -np.inf and np.all(grad[0] == 0)


def test_ragged_batch_many_pairs_waves(ctx):
    """More pairs than one wave of a deliberately small workspace: results independent of the wave split."""
    from gpbo_pkg import pkg

    t, y = orc.synthetic_trajectories(3, 130, seed=5)
    rng = np.random.default_rng(1)
    theta = np.log(np.array([1.0, 0.1, 1e-2]))[None, :] + 0.3 * rng.standard_normal((37, 3))
    gp_of = rng.integers(0, 3, 37).astype(np.int32)
    T = np.tile(t, (3, 1))
    big = ctx.lml_grad(T, y, theta, gp_of)
    small_ctx = pkg.Context(0, max_workspace_bytes=8 << 20)   # 8 MiB: a few pairs per wave
    small_ctx.set_small_path(ctx.small_max)
    assert small_ctx.wave_capacity(130) < 37
    small = small_ctx.lml_grad(T, y, theta, gp_of)
    small_ctx.close()
    for a, b in zip(big, small):
        assert np.array_equal(a, b)
    for k in (0, 17, 36):
        l0, g0, _ = orc.np_lml_grad(t, y[gp_of[k]], theta[k])
        assert abs(big[0][k] - l0) <= 1e-10 * abs(l0)


def test_replicated_pairs_are_bitwise_identical(ctx):
    """Race detector: 600 copies of one (GP, theta) pair spread over the SMs must give bit-identical results
    (all reductions are deterministic), equal to the oracle, with every status 0."""
    t, y = orc.synthetic_trajectories(1, 333, seed=5)
    th = np.log([1.9, 0.06, 4e-3])
    B = 600
    lml, grad, st = ctx.lml_grad(t[None], y, np.tile(th, (B, 1)), np.zeros(B, dtype=np.int32))
    assert np.all(st == 0)
    assert np.all(lml == lml[0]) and np.all(grad == grad[0])
    l0, g0, _ = orc.np_lml_grad(t, y[0], th)
    assert abs(lml[0] - l0) <= 1e-10 * abs(l0)
    assert rel(grad[0], g0, max(1.0, np.abs(g0).max())) <= 1e-9
    # same for the posterior-moment path
    t_est = np.linspace(0, 1, 150)
    G = 160
    state, ddt, cov, st2 = ctx.lstsq_moments(np.tile(t, (G, 1)), np.tile(y, (G, 1)), np.tile(th, (G, 1)), t_est)
    assert np.all(st2 == 0)
    assert np.all(state == state[0]) and np.all(ddt == ddt[0]) and np.all(cov == cov[0])


@pytest.mark.parametrize("m,B", [(4096, 3), (16384, 1)])
def test_full_size_gradient_matches_finite_differences(ctx, m, B):
    """BASELINE sizes (configs[3]: m = 4096, configs[4]: m = 16384), where the CPU oracle is too slow to run in a
    test: the analytic gradient (K^-1 trace reductions) must equal central differences of the LML itself --
    a size-independent consistency property between the Cholesky/solve path and the trtri/lauum path."""
    t, y = orc.synthetic_trajectories(B, m, seed=17)
    T = np.tile(t, (B, 1))
    th0 = np.log(np.array([[1.5, 0.05, 1e-2], [0.6, 0.02, 3e-2], [3.0, 0.1, 5e-3]])[:B])
    h = 1e-4
    thetas, gp_of = [th0], [np.arange(B)]
    for i in range(3):
        for sgn in (+1, -1):
            d = th0.copy()
            d[:, i] += sgn * h
            thetas.append(d)
            gp_of.append(np.arange(B))
    lml, grad, st = ctx.lml_grad(T, y, np.vstack(thetas), np.concatenate(gp_of).astype(np.int32))
    assert np.all(st == 0) and np.all(np.isfinite(lml)) and np.all(np.isfinite(grad))
    lml = lml.reshape(7, B)
    g = grad[:B]
    for i in range(3):
        fd = (lml[1 + 2 * i] - lml[2 + 2 * i]) / (2 * h)
        scale = np.maximum(1.0, np.abs(g).max(1))
        assert np.all(np.abs(fd - g[:, i]) <= 2e-3 * scale), (m, i, fd, g[:, i])


@pytest.mark.parametrize("m", [1, 2, 3, 17, 128, 129])
def test_tiny_and_boundary_sizes(ctx, m):
    """m = 1 ... and the sizes around one 128-tile; duplicate abscissae; extreme length scales."""
    rng = np.random.default_rng(m)
    t = np.sort(rng.uniform(0, 1, m))
    if m >= 3:
        t[1] = t[0]                                   # duplicate sample time: K is singular without the noise term
    y = np.sin(5 * t) + 0.1 * rng.standard_normal(m)
    thetas = np.log(np.array([[1.3, 0.2, 1e-2], [0.5, 1e-4, 1e-1], [2.0, 50.0, 1e-3]]))   # ell tiny / huge
    lml, grad, st = ctx.lml_grad(t[None], y[None], thetas, np.zeros(3, dtype=np.int32))
    for k in range(3):
        l0, g0, s0 = orc.np_lml_grad(t, y, thetas[k])
        assert st[k] == s0 == 0
        assert abs(lml[k] - l0) <= 1e-10 * max(1.0, abs(l0)), (m, k, lml[k], l0)
        assert rel(grad[k], g0, max(1.0, np.abs(g0).max())) <= 1e-8, (m, k, grad[k], g0)
    t_est = np.linspace(0, 1, 5)
    state, ddt, cov, w, st2, wst, _ = ctx.lstsq_weights(t[None], y[None], thetas[:1], t_est, 1e-8)
    ref = orc.np_lstsq_moments(t, y, thetas[0], t_est, want_sqrtW=False)
    assert st2[0] == 0 and wst[0] == 0
    assert rel(state[0], ref["state_estimate"]) <= 1e-10 and rel(cov[0], ref["ddt_covariance"]) <= 1e-9
    mean, std, alpha, _ = ctx.predict(t[None], y[None], thetas[:1], t_est[:1], want_alpha=True)
    m0, s0 = orc.np_predict(t, y, thetas[0], t_est[:1])
    assert rel(mean[0], m0, max(1e-300, np.abs(m0).max())) <= 1e-10 and abs(std[0, 0] - s0[0]) <= 1e-7 * max(1.0, s0[0])


# ------------------------------------------------------------------ assembly
@pytest.mark.parametrize("n1,n2", [(90, 90), (257, 33), (400, 200)])
def test_assemble_kinds(ctx, n1, n2):
    import torch

    rng = np.random.default_rng(0)
    t1 = np.sort(rng.uniform(0, 1, n1))
    t2 = t1 if n1 == n2 else np.sort(rng.uniform(0, 1, n2))
    theta = np.log(np.array([[1.7, 0.07, 3e-3], [0.4, 0.3, 1e-2]]))
    dev = torch.device("cuda", 0)
    a, b, th = (torch.as_tensor(x, device=dev) for x in (t1, t2, theta))
    out = torch.empty((2, n1, n2), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    for kind in range(7):
        if kind in (0, 1) and n1 != n2:
            continue
        ctx.assemble_device(kind, a.data_ptr(), 0, n1, b.data_ptr(), 0, n2, th.data_ptr(), 2, out.data_ptr(), 0)
        got = out.cpu().numpy()
        for p in range(2):
            s2, ell, chi = np.exp(theta[p])
            d = t1[:, None] - t2[None, :]
            kap = orc.np_rbf_eval(t1, t2, s2, ell)
            if kind == 0:
                ref = orc.np_kernel(t1, theta[p])
            elif kind == 1:
                ref = kap + np.diag(np.full(n1, chi))
            elif kind == 2:
                ref = orc.np_kernel(t1, theta[p], t2=t2)
            elif kind == 3:
                ref = kap
            elif kind == 4:
                ref = -d * kap / ell**2
            elif kind == 5:
                ref = (1 - (d**2 / ell**2)) * kap / ell**2
            else:
                dx2 = (t1[:, None] / ell - t2[None, :] / ell) ** 2
                ref = s2 * (np.exp(-0.5 * dx2) * dx2)
            assert rel(got[p], ref) <= 1e-14, (kind, p)


@pytest.mark.parametrize("n", [1, 63, 90, 257, 1000])
def test_assemble_symmetric_path(ctx, n):
    """t1 is t2 (same device pointer): the lower-triangle kernel that mirrors tiles through shared memory."""
    import torch

    rng = np.random.default_rng(n)
    t = np.sort(rng.uniform(0, 1, n))
    theta = np.log(np.array([[1.7, 0.07, 3e-3], [0.4, 0.004, 1e-2], [3.0, 1.5, 1e-6]]))   # incl. exp underflow
    dev = torch.device("cuda", 0)
    a, th = (torch.as_tensor(x, device=dev) for x in (t, theta))
    out = torch.empty((3, n, n), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    for kind in range(7):
        out.fill_(float("nan"))
        ctx.assemble_device(kind, a.data_ptr(), 0, n, a.data_ptr(), 0, n, th.data_ptr(), 3, out.data_ptr(), 0)
        got = out.cpu().numpy()
        for p in range(3):
            s2, ell, chi = np.exp(theta[p])
            d = t[:, None] - t[None, :]
            kap = orc.np_rbf_eval(t, t, s2, ell)
            dx2 = (t[:, None] / ell - t[None, :] / ell) ** 2
            ref = {0: lambda: orc.np_kernel(t, theta[p]), 1: lambda: kap + np.diag(np.full(n, chi)),
                   2: lambda: orc.np_kernel(t, theta[p], t2=t), 3: lambda: kap, 4: lambda: -d * kap / ell**2,
                   5: lambda: (1 - (d**2 / ell**2)) * kap / ell**2,
                   6: lambda: s2 * (np.exp(-0.5 * dx2) * dx2)}[kind]()
            assert np.all(np.isfinite(got[p]))
            # elementwise: 4 ulp relative, plus the flush-to-zero of results below 2^-1021 (fastmath.h); the cross
            # kinds 3-5 fold 1/ell^2 into constants, which perturbs the exponent argument by ~1 ulp -> |arg| ulp
            argmag = (d**2 / (2 * ell**2)) if kind in (3, 4, 5) else 0.0
            tol = (1e-15 + 4e-16 * argmag) * np.abs(ref) + 1e-300 * max(1.0, np.abs(ref).max())
            if kind == 5:
                tol = tol + 1e-15 * np.abs(kap) / ell**2      # cancellation in (1 - d^2/ell^2) near d = ell
            assert np.all(np.abs(got[p] - ref) <= tol), (kind, p)


@pytest.mark.parametrize("twice_nu", [3, 5])
@pytest.mark.parametrize("same", [False, True])
def test_assemble_matern(ctx, twice_nu, same):
    """Matern extension of the assembly (the reference has no Matern kernel): K and dK/dlog(ell) against
    scikit-learn's (ConstantKernel * Matern) + WhiteKernel, the derivative cross-covariances against the analytic
    formulas and against central differences of the kernel itself."""
    import torch
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel

    nu = twice_nu / 2
    rng = np.random.default_rng(twice_nu)
    n1, n2 = (150, 150) if same else (97, 61)
    t1 = np.sort(rng.uniform(0, 1, n1))
    t2 = t1 if same else np.sort(rng.uniform(0, 1, n2))
    s2, ell, chi = 1.7, 0.13, 4e-3
    theta = np.log(np.array([[s2, ell, chi]]))
    dev = torch.device("cuda", 0)
    a, th = torch.as_tensor(t1, device=dev), torch.as_tensor(theta, device=dev)
    b = a if same else torch.as_tensor(t2, device=dev)
    out = torch.empty((1, n1, n2), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    def run(kind):
        ctx.assemble_device(kind, a.data_ptr(), 0, n1, b.data_ptr(), 0, n2, th.data_ptr(), 1, out.data_ptr(), 0,
                            twice_nu=twice_nu)
        return out.cpu().numpy()[0].copy()

    kern = ConstantKernel(s2) * Matern(length_scale=ell, nu=nu) + WhiteKernel(chi)
    cross = kern(t1[:, None], t2[:, None])                        # White(X, Y) = 0
    assert rel(run(2), cross) <= 1e-14 and rel(run(3), cross) <= 1e-14
    if same:
        Kref, dK = kern(t1[:, None], eval_gradient=True)
        assert rel(run(0), Kref) <= 1e-14 and rel(run(1), Kref) <= 1e-14
        assert rel(run(6), dK[:, :, 1]) <= 1e-13              # theta order: constant, length scale, noise
    # derivative cross-covariances: analytic, and finite differences of the kernel
    tau = t1[:, None] - t2[None, :]
    aa = np.sqrt(twice_nu) / ell
    e = np.exp(-aa * np.abs(tau))
    if twice_nu == 3:
        k1 = -s2 * aa**2 * tau * e
        k2 = s2 * aa**2 * (1 - aa * np.abs(tau)) * e
    else:
        k1 = -s2 * aa**2 / 3 * tau * (1 + aa * np.abs(tau)) * e
        k2 = s2 * aa**2 / 3 * (1 + aa * np.abs(tau) - (aa * tau) ** 2) * e
    assert rel(run(4), k1) <= 1e-13 and rel(run(5), k2) <= 1e-13
    km = ConstantKernel(s2) * Matern(length_scale=ell, nu=nu)
    h = 1e-6
    fd1 = (km((t1 + h)[:, None], t2[:, None]) - km((t1 - h)[:, None], t2[:, None])) / (2 * h)
    mask = np.abs(tau) > 10 * h                                   # nu = 3/2 has a kink in dk/dtau at tau = 0
    assert np.abs(fd1 - k1)[mask].max() <= 1e-6 * np.abs(k1).max()


# ------------------------------------------------------------------ posterior moments
@pytest.mark.parametrize("name", REAL_CONFIGS)
def test_predict_and_lstsq_moments_golden(ctx, name):
    g = load_golden(name)
    T, Y, t_est, th = g["T"], g["Y"], g["t_est"], g["theta_opt"]
    mean, std, alpha, st = ctx.predict(T, Y, th, t_est, want_alpha=True)
    assert np.all(st == 0)
    state, ddt, cov, st2 = ctx.lstsq_moments(T, Y, th, t_est)
    assert np.all(st2 == 0)
    for gi in range(T.shape[0]):
        for key, got, ref in (("alpha", alpha[gi], g["alpha_opt"][gi]), ("pred_mean", mean[gi], g["pred_mean"][gi]),
                              ("pred_std", std[gi], g["pred_std"][gi]), ("state", state[gi], g["state_estimate"][gi]),
                              ("ddt", ddt[gi], g["ddt_estimate"][gi])):
            record(f"{key}_rel[{name}]", rel(got, ref))
        if "truth_alpha_opt" in g.files:
            ca, la = rel(alpha[gi], g["truth_alpha_opt"][gi]), rel(g["alpha_opt"][gi], g["truth_alpha_opt"][gi])
            record(f"alpha_vs_truth[{name}] cuda", ca)
            record(f"alpha_vs_truth[{name}] lapack", la)
            assert ca <= max(1e-10, 3 * la), (gi, ca, la)
        assert rel(alpha[gi], g["alpha_opt"][gi]) <= 1e-9
        assert rel(mean[gi], g["pred_mean"][gi]) <= 1e-10
        assert rel(std[gi], g["pred_std"][gi]) <= 1e-8      # sqrt of a difference of O(1) terms
        assert rel(state[gi], g["state_estimate"][gi]) <= 1e-10
        assert rel(ddt[gi], g["ddt_estimate"][gi]) <= 1e-10
    for gi in range(g["ddt_covariance"].shape[0]):
        record(f"ddt_cov_rel[{name}]", rel(cov[gi], g["ddt_covariance"][gi]))
        assert rel(cov[gi], g["ddt_covariance"][gi]) <= 1e-9
        assert np.array_equal(cov[gi], cov[gi].T)


def test_moments_against_oracle_tiled_sizes(ctx):
    """m and m' that span several 128-tiles and are not multiples of the tile edge."""
    t, y = orc.synthetic_trajectories(2, 300, seed=2)
    th = np.log(np.array([[2.5, 0.05, 1e-2], [0.7, 0.3, 3e-3]]))
    t_est = np.linspace(0, 1, 389)
    T = np.tile(t, (2, 1))
    state, ddt, cov, st = ctx.lstsq_moments(T, y, th, t_est)
    mean, std, alpha, _ = ctx.predict(T, y, th, t_est, want_alpha=True)
    for gi in range(2):
        ref = orc.np_lstsq_moments(t, y[gi], th[gi], t_est, want_sqrtW=False)
        assert rel(state[gi], ref["state_estimate"]) <= 1e-10
        assert rel(ddt[gi], ref["ddt_estimate"]) <= 1e-10
        assert rel(cov[gi], ref["ddt_covariance"]) <= 1e-9
        m0, s0 = orc.np_predict(t, y[gi], th[gi], t_est)
        assert rel(mean[gi], m0) <= 1e-10
        assert rel(std[gi], s0) <= 1e-7


def test_moments_many_estimation_points(ctx):
    """The m' >> m shape of `PDEs/experiments.sh` (m = 200 samples, m' = 3200 regression points), fitted Euler
    hyper-parameters, per-GP estimation grids."""
    g = load_golden("euler_006_200_03_400_6")
    T, Y, th = g["T"][:2], g["Y"][:2], g["theta_opt"][:2]
    t_est = np.stack([np.linspace(0, 0.06, 3200), np.linspace(0.001, 0.059, 3200)])
    state, ddt, cov, w, st, wst, _ = ctx.lstsq_weights(T, Y, th, t_est, 1e-8)
    assert np.all(st == 0) and np.all(wst == 0)
    for gi in range(2):
        ref = orc.np_lstsq_moments(T[gi], Y[gi], th[gi], t_est[gi], want_sqrtW=False)
        assert rel(state[gi], ref["state_estimate"]) <= 1e-9
        assert rel(ddt[gi], ref["ddt_estimate"]) <= 1e-9
        assert rel(cov[gi], ref["ddt_covariance"]) <= 1e-8
    # sqrtW through its defining identity, against the same residual of the reference's eigh route (cond ~ 1e11:
    # both are limited by cond * eps, gpkernels.py:496-504)
    x = np.random.default_rng(0).standard_normal(3200)
    A = cov[0] + 1e-8 * np.eye(3200)
    ours = np.abs(w[0] @ (A @ (w[0] @ x)) - x).max() / np.abs(x).max()
    wref, _ = _eigh_sqrtw(cov[0], 1e-8)
    theirs = np.abs(wref @ (A @ (wref @ x)) - x).max() / np.abs(x).max()
    record("sqrtw_identity_residual[m'=3200] newton-schulz", ours)
    record("sqrtw_identity_residual[m'=3200] eigh", theirs)
    assert ours <= max(1e-4, 3 * theirs), (ours, theirs)


# ------------------------------------------------------------------ optimiser
@pytest.mark.parametrize("name,tol", [("heat_1_20_05_80_5", 1e-8), ("seird_090_090_10_360", 1e-8),
                                      ("euler_006_200_03_400_6", 1e-8), ("seird_120_010_05_480", 1e-8),
                                      ("euler_006_050_01_400_6", 1e-8)])
def test_fit_reaches_reference_optimum(ctx, name, tol):
    """Same data, bounds and restart points as the reference run -> best LML within 1e-8 relative."""
    g = load_golden(name)
    T, Y = g["T"], g["Y"]
    G = T.shape[0]
    S = int(g["n_restarts"]) + 1
    starts = np.zeros((G, S, 3))
    starts[:, 1:] = g["starts"]
    gp_of = np.repeat(np.arange(G, dtype=np.int32), S)
    res = ctx.fit(T, Y, np.log(g["bounds"]), starts.reshape(-1, 3), gp_of)
    best = -np.nanmin(np.where(np.isfinite(res["fun"]), res["fun"], np.inf).reshape(G, S), axis=1)
    for gi in range(G):
        assert abs(best[gi] - g["lml_opt"][gi]) <= tol * abs(g["lml_opt"][gi]), (gi, best[gi], g["lml_opt"][gi])
    assert res["evals"] == int(res["nfev"].sum())


# ------------------------------------------------------------------ sqrtW = (C + eta I)^(-1/2)
def _eigh_sqrtw(C, eta):
    ev, V = np.linalg.eigh(C + eta * np.eye(C.shape[0]))       # gpkernels.py:496-504
    return V @ np.diag(1 / np.sqrt(ev)) @ V.T, ev


@pytest.mark.parametrize("name", REAL_CONFIGS)
def test_sqrtw_real_configs(ctx, name):
    """Newton-Schulz sqrtW on the reference's own derivative covariances: defining identity at least as tight as
    the reference's eigh route, element-wise agreement at the level the ill-conditioning allows."""
    g = load_golden(name)
    C = g["ddt_covariance"]
    eta = float(g["eta"])
    W, st, it = ctx.sqrtw(C, eta)
    assert np.all(st == 0) and np.all(it < 80)
    for k in range(C.shape[0]):
        n = C.shape[1]
        A = C[k] + eta * np.eye(n)
        Wref, ev = _eigh_sqrtw(C[k], eta)
        assert np.array_equal(W[k], W[k].T)
        res = np.abs(W[k] @ A @ W[k] - np.eye(n)).max()
        res_ref = np.abs(Wref @ A @ Wref - np.eye(n)).max()
        assert res <= max(3 * res_ref, 1e-9), (name, k, res, res_ref)
        cond = ev[-1] / ev[0]
        assert rel(W[k], Wref) <= max(1e-12, 1e-15 * cond * 10), (name, k, cond)


@pytest.mark.parametrize("n", [1, 80, 129, 300])
def test_sqrtw_well_conditioned_matches_eigh(ctx, n):
    rng = np.random.default_rng(n)
    M = rng.standard_normal((3, n, n))
    C = np.einsum("gij,gkj->gik", M, M) / n + 0.5 * np.eye(n)
    W, st, it = ctx.sqrtw(C, 1e-3)
    assert np.all(st == 0)
    for k in range(3):
        Wref, _ = _eigh_sqrtw(C[k], 1e-3)
        assert rel(W[k], Wref) <= 1e-12


def test_sqrtw_not_positive_definite_status_and_error(ctx, monkeypatch):
    n = 40
    rng = np.random.default_rng(0)
    M = rng.standard_normal((n, n))
    good = M @ M.T / n + np.eye(n)
    bad = good.copy()
    bad -= 3.0 * np.outer(M[:, 0], M[:, 0]) / (M[:, 0] @ M[:, 0]) * np.linalg.eigvalsh(good).max()   # one negative eigenvalue
    W, st, it = ctx.sqrtw(np.stack([good, bad, -np.eye(n)]), 1e-8)
    assert list(st) == [0, 1, 1]
    Wref, _ = _eigh_sqrtw(good, 1e-8)
    assert rel(W[0], Wref) <= 1e-12
    # the drop-in raises the reference's ValueError (gpkernels.py:500-503)
    from gpbo_pkg import pkg

    t, y = orc.synthetic_trajectories(1, 60, seed=2)
    gp = pkg.GP_RBFW((1e-5, 1e5), (1e-5, 1e2), (1e-16, 1e2), 2).fit(t, y[0])
    with pytest.raises(ValueError, match="increase eta"):
        gp.compute_lstsq_matrices(np.linspace(0, 1, 50), eta=-1e6)


def test_lstsq_weights_one_call_equals_two(ctx):
    t, y = orc.synthetic_trajectories(3, 200, seed=9)
    T = np.tile(t, (3, 1))
    theta = np.log(np.array([[1.5, 0.05, 1e-2], [0.8, 0.1, 3e-3], [2.0, 0.2, 1e-3]]))
    t_est = np.linspace(0, 1, 260)
    s1, d1, c1, st1 = ctx.lstsq_moments(T, y, theta, t_est)
    w1, wst1, _ = ctx.sqrtw(c1, 1e-8)
    s2, d2, c2, w2, st2, wst2, _ = ctx.lstsq_weights(T, y, theta, t_est, 1e-8)
    assert np.array_equal(s1, s2) and np.array_equal(d1, d2) and np.array_equal(c1, c2)
    assert np.array_equal(w1, w2) and np.array_equal(wst1, wst2)
    for k in range(3):
        A = c2[k] + 1e-8 * np.eye(260)
        assert np.abs(w2[k] @ A @ w2[k] - np.eye(260)).max() <= 1e-5


# ------------------------------------------------------------------ Matern extension (no counterpart in the reference)
@pytest.mark.parametrize("twice_nu", [3, 5])
@pytest.mark.parametrize("m", [90, 333, 700])
def test_matern_lml_grad_against_sklearn(ctx, twice_nu, m):
    t, y = orc.synthetic_trajectories(2, m, seed=m + twice_nu)
    thetas = np.log(np.array([[1.0, 0.1, 1e-3], [2.5, 0.05, 1e-2], [0.7, 0.6, 3e-3]]))
    T = np.tile(t, (2, 1))
    theta = np.tile(thetas, (2, 1))
    gp_of = np.repeat(np.arange(2, dtype=np.int32), len(thetas))
    ctx.set_kernel_family(twice_nu)
    try:
        lml, grad, st = ctx.lml_grad(T, y, theta, gp_of)
    finally:
        ctx.set_kernel_family(0)
    for k in range(len(theta)):
        l0, g0, s0 = orc.np_lml_grad_matern(t, y[gp_of[k]], theta[k], twice_nu)
        assert st[k] == s0 == 0
        assert abs(lml[k] - l0) <= 1e-10 * abs(l0), (twice_nu, m, k, lml[k], l0)
        assert rel(grad[k], g0, max(1.0, np.abs(g0).max())) <= 1e-9, (twice_nu, m, k, grad[k], g0)


@pytest.mark.parametrize("twice_nu", [3, 5])
def test_matern_moments_against_oracle(ctx, twice_nu):
    t, y = orc.synthetic_trajectories(2, 300, seed=40 + twice_nu)
    th = np.log(np.array([[2.5, 0.08, 1e-2], [0.7, 0.3, 3e-3]]))
    t_est = np.linspace(0, 1, 389)
    T = np.tile(t, (2, 1))
    ctx.set_kernel_family(twice_nu)
    try:
        state, ddt, cov, w, st, wst, _ = ctx.lstsq_weights(T, y, th, t_est, 1e-8)
        mean, std, alpha, _ = ctx.predict(T, y, th, t_est, want_alpha=True)
    finally:
        ctx.set_kernel_family(0)
    assert np.all(st == 0) and np.all(wst == 0)
    for gi in range(2):
        ref = orc.np_lstsq_moments_matern(t, y[gi], th[gi], t_est, twice_nu)
        assert rel(state[gi], ref["state_estimate"]) <= 1e-10
        assert rel(ddt[gi], ref["ddt_estimate"]) <= 1e-10
        assert rel(cov[gi], ref["ddt_covariance"]) <= 1e-9
        m0, s0, a0 = orc.np_predict_matern(t, y[gi], th[gi], t_est, twice_nu)
        assert rel(mean[gi], m0) <= 1e-10 and rel(std[gi], s0) <= 1e-7 and rel(alpha[gi], a0) <= 1e-9
        A = cov[gi] + 1e-8 * np.eye(389)
        assert np.abs(w[gi] @ A @ w[gi] - np.eye(389)).max() <= 1e-5


@pytest.mark.parametrize("nu", [1.5, 2.5])
def test_matern_fit_reaches_sklearn_optimum(nu):
    """GP_MaternW.fit against scikit-learn's own GaussianProcessRegressor with the Matern kernel, same restart points
    (global NumPy RNG, same seed): best LML within 1e-8 relative."""
    from gpbo_pkg import pkg

    t, y = orc.synthetic_trajectories(1, 80, seed=int(10 * nu))
    bounds = ((1e-5, 1e5), (1e-5, 1e2), (1e-16, 1e2))
    np.random.seed(99)
    ref = orc.OracleGP(*bounds, 12, twice_nu=int(2 * nu)).fit(t, y[0])
    np.random.seed(99)
    gp = pkg.GP_MaternW(nu, *bounds, 12).fit(t, y[0])
    assert abs(gp.gpr.log_marginal_likelihood_value_ - ref.lml) <= 1e-8 * abs(ref.lml)
    ts = np.linspace(0, 1, 57)
    mean, std = gp.predict(ts)
    m0, s0, _ = orc.np_predict_matern(t, y[0], gp.gpr.kernel_.theta, ts, int(2 * nu))
    assert rel(mean, m0) <= 1e-10 and rel(std, s0) <= 1e-7
    gp.compute_lstsq_matrices(ts, eta=1e-8)
    refm = orc.np_lstsq_moments_matern(t, y[0], gp.gpr.kernel_.theta, ts, int(2 * nu))
    assert rel(gp.ddt_estimate, refm["ddt_estimate"]) <= 1e-10 and rel(gp.ddt_covariance, refm["ddt_covariance"]) <= 1e-9
    assert "Matern" in str(gp) and rel(gp.rbf_eval(ts, t), orc.np_matern(ts, t, gp.constant, gp.length_scale, int(2 * nu))) <= 1e-13
    # the batched step keeps working with the extension kernel and leaves RBF objects unaffected
    gps = pkg.fit_gaussian_processes(ts, t, y, 1e-8, constant_bounds=bounds[0], length_scale_bounds=bounds[1],
                                     noise_level_bounds=bounds[2], n_restarts_optimizer=3, verbose=False,
                                     kernel="matern32" if nu == 1.5 else "matern52")
    assert isinstance(gps[0], pkg.GP_MaternW) and gps[0].sqrtW.shape == (57, 57)
    rbf = pkg.GP_RBFW(*bounds, 2).fit(t, y[0])
    l0, g0, _ = orc.np_lml_grad(t, y[0], rbf.gpr.kernel_.theta)
    assert abs(rbf.gpr.log_marginal_likelihood(rbf.gpr.kernel_.theta) - l0) <= 1e-10 * abs(l0)


def test_weighted_products(ctx):
    """sqrtW_i @ D and sqrtW_i @ z_i of WeightedLSTSQSolver.fit (wlstsq.py:183-188), with the weights passed from the
    host and with the weights left resident on the device by lstsq_weights."""
    t, y = orc.synthetic_trajectories(3, 150, seed=21)
    T = np.tile(t, (3, 1))
    theta = np.log(np.array([[1.5, 0.05, 1e-2], [0.8, 0.1, 3e-3], [2.0, 0.2, 1e-3]]))
    t_est = np.linspace(0, 1, 203)
    state, ddt, cov, w, st, wst, _ = ctx.lstsq_weights(T, y, theta, t_est, 1e-6)
    rng = np.random.default_rng(0)
    D = rng.standard_normal((203, 27))                      # d not a multiple of the column chunk
    ol, orr = ctx.weighted_products(D, ddt)                 # resident weights
    for g in range(3):
        # summation order differs from BLAS: compare against the rounding scale |W| |x| (sqrtW whitens a smooth
        # ddt estimate, so the product itself is small by cancellation)
        assert np.all(np.abs(ol[g] - w[g] @ D) <= 1e-14 * (np.abs(w[g]) @ np.abs(D)))
        assert np.all(np.abs(orr[g] - w[g] @ ddt[g]) <= 1e-14 * (np.abs(w[g]) @ np.abs(ddt[g])))
    ol2, or2 = ctx.weighted_products(D[:, :1], ddt, sqrtW=w)
    assert np.array_equal(ol2[:, :, 0], ol[:, :, 0]) and np.array_equal(or2, orr)
    with pytest.raises(Exception, match="another shape"):
        ctx.weighted_products(D[:100], ddt[:, :100])
