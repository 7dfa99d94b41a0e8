"""Scratch GPU probe: latency of one batched LML+grad evaluation with few pairs in flight (the optimiser's tail) and a
bounded fit to convergence (bench.fit_sample), for A/B runs of GPBO_OVERLAP_MAX."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench
from gpbo_pkg import pkg

m = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = pkg.default_context(0)
T, Y, theta, gp_of = pkg.workload.eval_workload(4, m, 32)
ctx.upload_problem(T, Y)
out = {"m": m, "overlap_max": os.environ.get("GPBO_OVERLAP_MAX", "default"), "ms_per_eval_call": {}}
BATCHES = tuple(int(x) for x in os.environ.get("TAIL_BATCHES", "1,2,4,8,16,32,48,64,96").split(","))
for B in BATCHES:
    ctx.lml_grad_resident(theta[:B], gp_of[:B])
    t0 = time.perf_counter()
    for _ in range(5):
        ctx.lml_grad_resident(theta[:B], gp_of[:B])
    out["ms_per_eval_call"][B] = round((time.perf_counter() - t0) / 5 * 1e3, 3)
ctx.profile_enable(True)
ctx.lml_grad_resident(theta[:1], gp_of[:1])
out["profile_ms_B1"] = {k: round(v[0], 3) for k, v in ctx.profile_get().items() if v[1]}
ctx.profile_enable(False)
if len(sys.argv) > 2:
    fs = bench.fit_sample(pkg, ctx, r=4, m=m, S=32)
    out["fit_sample"] = {k: fs[k] for k in ("fit_seconds", "lml_grad_evals", "rounds", "evals_per_s_over_fit")}
    out["fit_sample"]["best_lml"] = fs["best_lml"]
print(json.dumps(out))
