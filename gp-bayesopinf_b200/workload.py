"""Synthetic step2_fitgps workloads of BASELINE.json's configs[3] / configs[4].

The reference ships data generators only for its three experiments (full-order ODE / PDE solves); the synthetic sizes
(r = 64 x n = 4096 / 8192, r = 256 x n = 16384) use the signal proposed in the survey: sorted uniform sample times
with the end points forced (like ``PDEs/step1_generate_data.py:48-56``), three sinusoids per mode and 3 % noise;
hyper-parameter bounds and restart semantics are the Euler configuration's (``PDEs/config_euler.py:100-103``,
sklearn ``_gpr.py:312-333``: start 0 at theta = log(1, 1, 1), the others log-uniform in the box).
"""

from __future__ import annotations

import numpy as np

EULER_BOUNDS = np.array([(1e-5, 1e5), (1e-5, 1e2), (1e-16, 1e2)])      # PDEs/config_euler.py:100-102


def synthetic_trajectories(r, m, seed=0):
    """t = sort(U(0,1)) with endpoints forced, y_i = sum_k a sin(2 pi f t + phi) + 0.03 N(0,1).  -> t (m,), y (r, m)."""
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0.0, 1.0, size=m))
    t[0], t[-1] = 0.0, 1.0
    a = rng.uniform(0.5, 2.0, size=(r, 3))
    f = rng.uniform(1.0, 8.0, size=(r, 3))
    ph = rng.uniform(0.0, 2 * np.pi, size=(r, 3))
    y = (a[:, :, None] * np.sin(2 * np.pi * f[:, :, None] * t[None, None, :] + ph[:, :, None])).sum(1)
    y += 0.03 * rng.standard_normal((r, m))
    return t, y


def fit_workload(r, m, S, seed=0, bounds=EULER_BOUNDS):
    """(T (r, m), Y (r, m), bounds_log (3, 2), starts (r * S, 3), gp_of (r * S,)) of a multi-start fit with S starts
    per mode: start 0 at theta = 0 like sklearn, the other S - 1 log-uniform in the box (sklearn ``_gpr.py:328-333``)."""
    t, y = synthetic_trajectories(r, m, seed=seed)
    bl = np.log(np.asarray(bounds, dtype=np.float64))
    rng = np.random.default_rng(seed + 7)
    starts = rng.uniform(bl[:, 0], bl[:, 1], size=(r, S, 3))
    starts[:, 0] = 0.0
    gp_of = np.repeat(np.arange(r, dtype=np.int32), S)
    return np.tile(t, (r, 1)), y, bl, np.ascontiguousarray(starts.reshape(-1, 3)), gp_of


def eval_workload(r, m, S, seed=0):
    """One theta per (mode, start) pair, log-uniform in the part of the Euler box where the optimiser spends its time
    (all K positive definite): the fixed-theta batch of the LML + gradient throughput measurement."""
    t, y = synthetic_trajectories(r, m, seed=seed)
    rng = np.random.default_rng(seed + 1)
    lo = np.log([0.3, 0.01, 1e-4])
    hi = np.log([10.0, 0.2, 1e-1])
    theta = rng.uniform(lo, hi, size=(r * S, 3))
    gp_of = np.repeat(np.arange(r, dtype=np.int32), S)
    return np.tile(t, (r, 1)), y, theta, gp_of
