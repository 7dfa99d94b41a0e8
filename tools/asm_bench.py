"""Scratch GPU probe: achieved HBM GB/s of the stand-alone kernel-matrix assembly (gpbo_assemble) and timing of
the posterior-moment path, against MEASURED_PEAKS.json's copy bandwidth."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
from gpbo_pkg import pkg
import gp_oracle as orc

ctx = pkg.default_context(0)
dev = torch.device("cuda", 0)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
out = {"hbm_peak_gbs": peak, "assemble": [], "moments": []}
stream = torch.cuda.current_stream(dev)
for n1, n2, B in ((4096, 4096, 32), (8192, 8192, 8), (16384, 16384, 2), (3200, 200, 64)):
    t1 = torch.sort(torch.rand(B, n1, dtype=torch.float64, device=dev), dim=1).values.contiguous()
    t2 = torch.sort(torch.rand(B, n2, dtype=torch.float64, device=dev), dim=1).values.contiguous()
    th = torch.log(torch.tensor([[1.5, 0.05, 1e-2]], dtype=torch.float64, device=dev)).repeat(B, 1).contiguous()
    o = torch.empty((B, n1, n2), dtype=torch.float64, device=dev)
    for kind, sym in [(k, s) for k in (0, 1, 2, 3, 4, 5, 6) for s in (False, True)]:
        if (kind in (0, 1) or sym) and n1 != n2:
            continue
        b2 = t1 if sym else t2
        args = (kind, t1.data_ptr(), n1, n1, b2.data_ptr(), n2, n2, th.data_ptr(), B, o.data_ptr(), stream.cuda_stream)
        ctx.assemble_device(*args)
        reps = 5
        torch.cuda.synchronize()
        ctx.profile_enable(True)            # kernel time from the library's own events (the call syncs the stream)
        for _ in range(reps):
            ctx.assemble_device(*args)
        pms, pn = ctx.profile_get()["assemble"]
        ctx.profile_enable(False)
        ms = pms / max(pn, 1)
        gbs = B * n1 * n2 * 8 / (ms * 1e-3) / 1e9
        out["assemble"].append({"kind": kind, "sym": sym, "n1": n1, "n2": n2, "B": B, "ms": ms, "GBps": gbs, "frac": gbs / peak,
                                "Gelem_s": B * n1 * n2 / (ms * 1e-3) / 1e9})
    del t1, t2, o
    torch.cuda.empty_cache()

for (G, m, n) in ((16, 4096, 4096), (6, 200, 3200), (25, 20, 80), (5, 90, 360)):
    t, y = orc.synthetic_trajectories(G, m, seed=3)
    T = np.tile(t, (G, 1))
    theta = np.tile(np.log([1.5, 0.05, 1e-2]), (G, 1))
    t_est = np.linspace(0, 1, n)
    ctx.lstsq_moments(T, y, theta, t_est)
    ctx.profile_enable(True)
    t0 = time.perf_counter()
    ctx.lstsq_moments(T, y, theta, t_est)
    dt = time.perf_counter() - t0
    prof = ctx.profile_get()
    ctx.profile_enable(False)
    t0 = time.perf_counter()
    ctx.predict(T, y, theta, t_est)
    dp = time.perf_counter() - t0
    out["moments"].append({"G": G, "m": m, "n_est": n, "lstsq_host_s": dt, "predict_host_s": dp,
                           "kernel_ms": {k: round(v[0], 3) for k, v in prof.items() if v[1]}})
print(json.dumps(out))
