"""Phase timing of the in-shared small-matrix kernels (run with GPBO_SMALL_DBG=1): fixed-theta batches and the
persistent fit on the reference configurations."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpbo_pkg import pkg  # noqa: E402

ctx = pkg.default_context(0)
for m in (20, 64, 90, 200, 224):
    t, y = pkg.workload.synthetic_trajectories(1, m, seed=m)
    th = np.tile(np.log([1.5, 0.1, 1e-2]), (1, 1))
    ctx.lml_grad(t[None], y, th)
    t0 = time.perf_counter()
    for _ in range(20):
        ctx.lml_grad(t[None], y, th)
    print(f"m={m}: single pair lml_grad call {1e6 * (time.perf_counter() - t0) / 20:.1f} us", flush=True)
    B = 148 * 4
    th = np.tile(np.log([1.5, 0.1, 1e-2]), (B, 1))
    ctx.lml_grad(t[None], y, th, np.zeros(B, dtype=np.int32))
    t0 = time.perf_counter()
    ctx.lml_grad(t[None], y, th, np.zeros(B, dtype=np.int32))
    print(f"m={m}: {B} pairs lml_grad call {1e3 * (time.perf_counter() - t0):.3f} ms", flush=True)
for name in ("heat_1_20_05_80_5", "seird_090_090_10_360", "euler_006_200_03_400_6"):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"), allow_pickle=False)
    T, Y = g["T"], g["Y"]
    G = T.shape[0]
    S = g["starts"].shape[1] + 1
    starts = np.zeros((G, S, 3))
    starts[:, 1:] = g["starts"]
    gp_of = np.repeat(np.arange(G, dtype=np.int32), S)
    for rep in range(2):
        t0 = time.perf_counter()
        res = ctx.fit(T, Y, np.log(g["bounds"]), starts.reshape(-1, 3), gp_of)
        dt = time.perf_counter() - t0
    nf = np.sort(res["nfev"])[::-1]
    print(f"{name}: fit {dt:.4f} s, evals {res['evals']}, longest chains {nf[:6].tolist()}, status histogram "
          f"{np.bincount(res['status'], minlength=6).tolist()}", flush=True)
