"""Add the 80-bit long-double truth (oracle/lml_ld.c) to the real-configuration goldens.

TEST INFRASTRUCTURE ONLY.  ``python oracle/add_truth.py`` augments ``tests/golden/<config>.npz`` (written by
``make_golden.py`` from the unmodified reference) with ``truth_lml_eval`` / ``truth_grad_eval`` for every fixed-theta
evaluation point and ``truth_alpha_opt`` at the fitted theta, so that GPU tests can ask whether the CUDA result is as
close to the exact value as the reference's FP64 LAPACK path is.  Needs ``oracle/_build/lml_ld`` (``make -C oracle``).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gp_oracle as orc  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")
CONFIGS = ("seird_090_090_10_360", "heat_1_20_05_80_5", "euler_006_200_03_400_6", "seird_120_010_05_480",
           "euler_006_050_01_400_6")

if __name__ == "__main__":
    for name in sys.argv[1:] or CONFIGS:
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        g = dict(np.load(path, allow_pickle=False))
        T, Y = g["T"], g["Y"]
        G, P = g["thetas_eval"].shape[:2]
        tl = np.full((G, P), np.nan)
        tg = np.full((G, P, 3), np.nan)
        ta = np.zeros_like(g["alpha_opt"])
        for gi in range(G):
            for k in range(P):
                if np.isfinite(g["lml_eval"][gi, k]):
                    tl[gi, k], tg[gi, k], _ = orc.ld_truth(T[gi], Y[gi], g["thetas_eval"][gi, k])
            _, _, ta[gi] = orc.ld_truth(T[gi], Y[gi], g["theta_opt"][gi])
        g["truth_lml_eval"], g["truth_grad_eval"], g["truth_alpha_opt"] = tl, tg, ta
        np.savez_compressed(path, **g)
        ok = np.isfinite(tl) & (g["cond_eval"] <= 1e6)
        el = np.abs(tl - g["lml_eval"])[ok] / np.maximum(1.0, np.abs(tl[ok]))
        print(f"{name}: {ok.sum()} points, reference vs truth: lml {el.max():.2e}", flush=True)
