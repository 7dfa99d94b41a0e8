"""Sharded step2 on N GPUs (run under torchrun): the Heat reference configuration (25 GPs x 101 starts) fitted with the
(GP x start) pairs sharded cyclically over the ranks and all-gathered over NCCL, checked on every rank against the
reference's golden optimum and against a single-rank run of the same call."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpbo_pkg import pkg

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = np.load(os.path.join(ROOT, "tests", "golden", "heat_1_20_05_80_5.npz"), allow_pickle=False)
T, Y, t_est, b, eta = g["T"], g["Y"], g["t_est"], g["bounds"], float(g["eta"])
L, r = 5, 5                                               # 5 trajectories x 5 modes
starts = iter(g["starts"])
pkg.step2_fitgps.draw_restart_points = lambda bl, n: next(starts)      # replay the reference's restart points
kw = dict(constant_bounds=tuple(b[0]), length_scale_bounds=tuple(b[1]), noise_level_bounds=tuple(b[2]),
          n_restarts_optimizer=int(g["n_restarts"]), verbose=False)
t0 = time.perf_counter()
gps = pkg.fit_gaussian_processes_multi(t_est, [T[l * r] for l in range(L)], [Y[l * r:(l + 1) * r] for l in range(L)], eta,
                                       group=True, **kw)
dt = time.perf_counter() - t0
flat = [gp for traj in gps for gp in traj]
lml = np.array([gp.gpr.log_marginal_likelihood_value_ for gp in flat])
gap = float(np.max((g["lml_opt"] - lml) / np.abs(g["lml_opt"])))
owned = [i for i, gp in enumerate(flat) if hasattr(gp, "sqrtW")]
ok_owned = owned == list(range(rank, len(flat), world))
state_err = float(max(np.abs(gp.state_estimate - g["state_estimate"][i]).max() / np.abs(g["state_estimate"][i]).max()
                      for i, gp in enumerate(flat)))
res_w = float(max(np.abs(flat[i].sqrtW @ (flat[i].ddt_covariance + eta * np.eye(t_est.size)) @ flat[i].sqrtW
                         - np.eye(t_est.size)).max() for i in owned))
th = torch.tensor([gp.gpr.kernel_.theta for gp in flat], device="cuda")
ths = [torch.empty_like(th) for _ in range(world)]
dist.all_gather(ths, th)
same = all(bool(torch.equal(ths[0], x)) for x in ths)
out = {"rank": rank, "world": world, "seconds": dt, "max_rel_lml_gap_vs_reference": gap, "covariances_owned_cyclically": ok_owned,
       "state_estimate_rel_err_vs_reference": state_err, "sqrtw_identity_residual": res_w, "theta_identical_on_all_ranks": same}
assert gap <= 1e-8 and ok_owned and same and state_err <= 1e-3 and res_w <= 1e-4, out
if rank == 0:
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
