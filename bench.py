#!/usr/bin/env python
"""bench.py -- headline benchmark of the step2_fitgps hot path (contract: see the task prompt / DESIGN.md §6).

Workload (BASELINE.json configs[3]): synthetic r=64 modes x n=4096 time points x 32 hyper-parameter starts
on one B200.  A *step* is one lock-step pass of the optimiser's inner evaluation over the whole batch:
LML + gradient (Cholesky + K^-1 traces) for all 64 x 32 = 2048 (mode, start) pairs at their current theta.
`value` = LML+grad evaluations per second, inputs resident in HBM.  `e2e` = the same through the host-pointer
C-ABI call (gpbo_lml_grad_host via the ctypes layer) with pinned host buffers: H2D of (t, y, theta, gp_of) and
D2H of (lml, grad, status) inside the timed region.  With N > 1 ranks every rank runs its own 64 x 32 batch
(weak scaling; modes shard with no data-path collective) and all-gathers the 2048 x 4 results over NCCL.

`--impl reference` times the reference's own CPU path for the same evaluation
(GaussianProcessRegressor.log_marginal_likelihood(theta, eval_gradient=True) driven exactly as
codebase/gpkernels.py does, through the oracle port) on the host cores, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R_MODES, N_POINTS, N_STARTS = 64, 4096, 32
METRIC, UNIT = "gp_lml_grad_evals_per_sec", "evals/s"


def workload(r, m, S, seed=0):
    """Synthetic trajectories and one theta per (mode, start) pair, drawn log-uniformly from
    the part of the Euler hyper-parameter box where the optimiser spends its time (all K positive definite)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from gp_oracle import synthetic_trajectories  # input generator shared with the golden fixtures

    t, y = synthetic_trajectories(r, m, seed=seed)
    rng = np.random.default_rng(seed + 1)
    lo = np.log([0.3, 0.01, 1e-4])
    hi = np.log([10.0, 0.2, 1e-1])
    theta = rng.uniform(lo, hi, size=(r * S, 3))
    gp_of = np.repeat(np.arange(r, dtype=np.int32), S)
    return np.tile(t, (r, 1)), y, theta, gp_of


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx = float(p[1])
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


def run_reference(args):
    """Reference arm: the reference's CPU evaluation of the same quantity, on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gp_oracle as orc
    from threadpoolctl import threadpool_info, threadpool_limits

    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core the BLAS can take
    threadpool_limits(limits=os.cpu_count())
    T, Y, theta, gp_of = workload(R_MODES, N_POINTS, N_STARTS)
    b = np.array([(1e-5, 1e5), (1e-5, 1e2), (1e-16, 1e2)])
    gp = orc.OracleGP(tuple(b[0]), tuple(b[1]), tuple(b[2]), 0)
    gp.gpr.optimizer = None
    gp.fit(T[0], Y[0])          # only stores the training data (no optimisation)
    times = []
    for it in range(args.warmup + args.steps):
        k = it % theta.shape[0]
        t0 = time.perf_counter()
        gp.lml_grad(theta[k])   # one pair of the 2048-pair step
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    per = sum(times) / len(times)
    cores = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    val = 1.0 / per
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"synthetic r={R_MODES} x n={N_POINTS} x {N_STARTS} starts (BASELINE configs[3])",
                   "sample": "1 of the 2048 (mode, start) LML+grad evaluations per step"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "sklearn log_marginal_likelihood(theta, eval_gradient=True), m=4096, one pair per step"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--modes", type=int, default=R_MODES)
    ap.add_argument("--points", type=int, default=N_POINTS)
    ap.add_argument("--starts", type=int, default=N_STARTS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fit-sample", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch

    from gpbo_pkg import pkg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    r, m, S = args.modes, args.points, args.starts
    T, Y, theta, gp_of = workload(r, m, S, seed=rank)      # every rank: its own r modes (weak scaling)
    B = theta.shape[0]
    ctx = pkg.default_context(local)

    # device-resident inputs/outputs (torch = allocator + address carrier)
    Td, Yd, thd = (torch.as_tensor(x, device=dev) for x in (T, Y, theta))
    gpd = torch.as_tensor(gp_of, device=dev)
    res = torch.empty((B, 4), dtype=torch.float64, device=dev)      # [lml, grad(3)] per pair
    lml_d, grad_d = torch.empty(B, dtype=torch.float64, device=dev), torch.empty((B, 3), dtype=torch.float64, device=dev)
    st_d = torch.empty(B, dtype=torch.int32, device=dev)
    gathered = torch.empty((world * B, 4), dtype=torch.float64, device=dev) if world > 1 else None
    stream = torch.cuda.current_stream(dev)

    def step_device():
        ctx.lml_grad_device(Td.data_ptr(), Yd.data_ptr(), r, m, thd.data_ptr(), gpd.data_ptr(), B, lml_d.data_ptr(),
                            grad_d.data_ptr(), st_d.data_ptr(), stream.cuda_stream)
        if world > 1:
            res[:, 0] = lml_d
            res[:, 1:] = grad_d
            dist.all_gather_into_tensor(gathered, res)

    # pinned host buffers for the end-to-end arm
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    Th, Yh, thh, gph = pin(T), pin(Y), pin(theta), pin(gp_of)

    def step_host():
        return ctx.lml_grad(Th, Yh, thh, gph)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        w0 = time.perf_counter()
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - w0
        ms = max(e0.elapsed_time(e1), 0.0)
        # the C ABI synchronises its stream before returning, so event time == wall time up to launch overheads;
        # take the max so copies issued on the library's own stream (host arm) are covered too
        ms = max(ms, wall * 1e3) if fn is step_host else ms
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms

    # FP64 tensor-pipe peak (DMMA issue rate), measured live: MEASURED_PEAKS.json has no FP64 entry
    dmma_peak, _ = ctx.dmma_peak(100000)

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.launch_count
    ctx.profile_enable(True)
    ms = timed(step_device, args.steps)
    prof = ctx.profile_get()
    ctx.profile_enable(False)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop()
    ok = int((st_d == 0).sum().item())
    ms_per_step = ms / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # end-to-end through the host-pointer C-ABI call
    step_host()
    ms_h = timed(step_host, max(1, min(args.steps, 2))) / max(1, min(args.steps, 2))
    e2e = world * B / (ms_h * 1e-3)
    h2d = Th.nbytes + Yh.nbytes + thh.nbytes + gph.nbytes
    d2h = B * (8 + 24 + 4)

    # roofline of the dominant kernel class, per launch: algorithmic flops = m_pad^3/3 per pair for each
    # of potrf (diag+panel), trtri, lauum; "issued" = DMMA flops the 128-tile schedule actually executes.
    T_blocks = (m + 127) // 128
    m_pad = T_blocks * 128
    blk = 2.0 * 128 ** 3
    # structural-zero skipping (tile_engine.cuh): symmetric diagonal tiles issue 136/256 of their 8x8 blocks; a k-block
    # with a triangular block inverse as row operand issues 36/64, as column operand 40/64 of its DMMAs
    SYM, TRI_A, TRI_B = 136.0 / 256.0, 36.0 / 64.0, 40.0 / 64.0
    issued_pp = {
        "chol_diag": blk * SYM * sum(j for j in range(T_blocks)),
        "chol_panel": blk * sum((T_blocks - 1 - j) * (j + TRI_B) for j in range(T_blocks)),
        "trtri": blk * sum((i - j - 1) + TRI_B + TRI_A for i in range(1, T_blocks) for j in range(i)),
        "lauum_grad": blk * sum(i * ((T_blocks - i - 1) + TRI_A) + SYM * (T_blocks - i) for i in range(T_blocks)),
    }
    dom = max(("chol_panel", "trtri", "lauum_grad"), key=lambda k: prof[k][0])
    cap = min(B, ctx.wave_capacity(m))
    waves = [min(cap, B - w0) for w0 in range(0, B, cap)]
    flops_third = float(m_pad) ** 3 / 3.0
    launches_dom = max(prof[dom][1], 1)
    total_flops_dom = flops_third * B * args.steps   # all launches of that class over the timed region
    dom_ms = prof[dom][0] + (prof["chol_diag"][0] if dom == "chol_panel" else 0.0)   # potrf = diag + panel
    achieved = total_flops_dom / (dom_ms * 1e-3) / 1e12
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(dom)
    except Exception:
        pass
    roofline = {
        "bound": "tensor", "kernel": dom, "achieved": achieved, "peak": dmma_peak, "unit": "TFLOP/s",
        "frac": achieved / dmma_peak, "traffic": traffic,
        "peak_source": "FP64 DMMA issue-rate micro-benchmark run in this process (gpbo_bench_dmma_peak); "
                       "MEASURED_PEAKS.json holds no FP64 figure",
        "flops_per_launch": total_flops_dom / launches_dom, "avg_launch_ms": dom_ms / launches_dom,
        "share_of_step": prof[dom][0] / sum(v[0] for v in prof.values()),
        "whole_eval_tflops": world * B * float(m) ** 3 / (ms_per_step * 1e-3) / 1e12,
        "whole_eval_frac": B * float(m) ** 3 / (ms_per_step * 1e-3) / 1e12 / dmma_peak,
        "dmma_issued_tflops": {k: issued_pp[k] * B * args.steps / (prof[k][0] * 1e-3) / 1e12
                               for k in issued_pp if prof[k][0] > 0},
        "kernel_ms": {k: round(v[0], 3) for k, v in prof.items() if v[1]},
        "waves": waves,
    }

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"synthetic r={r} modes x n={m} points x {S} starts per GPU (BASELINE configs[3]); "
                               "step = LML+gradient of all r*S (mode,start) pairs",
                   "pairs_per_gpu": B, "l2": "inputs larger than L2 (each pair's 128 MiB factor streams from HBM)",
                   "status_ok_pairs": ok},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": ms_h},
        "roofline": roofline,
    }

    if rank == 0 and world == 1 and not args.no_fit_sample:
        out["roofline_assembly"] = assembly_roofline(ctx, dev)
        out["fit_sample"] = fit_sample(ctx, m)
        out["fit_reference_configs"] = fit_reference_configs(ctx)
        out["moments_sample"] = moments_sample(ctx, m)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(T, Y, theta)
    elif rank == 0:
        out["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def fit_sample(ctx, m, r=4, S=32):
    """Measured GP fits/s on a bounded sub-workload: r modes x S starts, full L-BFGS-B + posterior moments."""
    T, Y, _, gp_of = workload(r, m, S, seed=123)
    b = np.log(np.array([(1e-5, 1e5), (1e-5, 1e2), (1e-16, 1e2)]))
    rng = np.random.default_rng(7)
    starts = rng.uniform(b[:, 0], b[:, 1], size=(r * S, 3))
    starts[::S] = 0.0
    t0 = time.perf_counter()
    res = ctx.fit(T, Y, b, starts, gp_of)
    t_fit = time.perf_counter() - t0
    funs = np.where(np.isfinite(res["fun"]), res["fun"], np.inf).reshape(r, S)
    best = res["theta"].reshape(r, S, 3)[np.arange(r), funs.argmin(1)]
    t1 = time.perf_counter()
    ctx.lstsq_moments(T, Y, best, np.linspace(0, 1, m))
    ctx.predict(T, Y, best, np.linspace(0, 1, m))
    t_mom = time.perf_counter() - t1
    return {"modes": r, "starts": S, "m": m, "m_est": m, "fit_seconds": t_fit, "moments_seconds": t_mom,
            "fits_per_s": r / (t_fit + t_mom), "lml_grad_evals": res["evals"], "rounds": res["rounds"],
            "best_lml": (-funs.min(1)).tolist()}


def assembly_roofline(ctx, dev, n=8192, B=8):
    """HBM-write roofline of the stand-alone kernel-matrix assembly (train K, sklearn order): 8 B per element."""
    import torch

    peaks = measured_peaks() or {}
    peak = float(peaks.get("hbm_gbs", 6543.1))
    t = torch.sort(torch.rand(B, n, dtype=torch.float64, device=dev), dim=1).values.contiguous()
    th = torch.log(torch.tensor([[1.5, 0.05, 1e-2]], dtype=torch.float64, device=dev)).repeat(B, 1).contiguous()
    o = torch.empty((B, n, n), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev)
    res = {}
    for name, kind, other in (("train_K_symmetric", 0, t), ("cross_K_general", 2, t.clone())):
        args = (kind, t.data_ptr(), n, n, other.data_ptr(), n, n, th.data_ptr(), B, o.data_ptr(), stream.cuda_stream)
        ctx.assemble_device(*args)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record(stream)
        for _ in range(5):
            ctx.assemble_device(*args)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        gbs = 5 * B * n * n * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9
        res[name] = {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                     "bytes_per_launch": B * n * n * 8, "shape": [B, n, n]}
    res["peak_source"] = "MEASURED_PEAKS.json hbm_gbs (burst copy bandwidth)" if peaks else "fallback 6543.1 GB/s"
    return res


def fit_reference_configs(ctx):
    """Full step2 fit (all 101 starts per GP, the reference's own start points) + posterior moments on five of the
    reference's experiment configurations (tests/golden/*.npz), with the reference's CPU wall time recorded when the
    fixtures were generated (8-core build container) beside it."""
    out = {}
    for name in ("seird_090_090_10_360", "seird_120_010_05_480", "heat_1_20_05_80_5", "euler_006_200_03_400_6",
                 "euler_006_050_01_400_6"):
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"), allow_pickle=False)
        T, Y, t_est = g["T"], g["Y"], g["t_est"]
        G = T.shape[0]
        bl = np.log(g["bounds"])
        S = g["starts"].shape[1] + 1
        starts = np.zeros((G, S, 3))
        starts[:, 1:] = g["starts"]
        gp_of = np.repeat(np.arange(G, dtype=np.int32), S)
        t0 = time.perf_counter()
        res = ctx.fit(T, Y, bl, starts.reshape(-1, 3), gp_of)
        funs = np.where(np.isfinite(res["fun"]), res["fun"], np.inf).reshape(G, S)
        best = res["theta"].reshape(G, S, 3)[np.arange(G), funs.argmin(1)]
        ctx.lstsq_moments(T, Y, best, t_est)
        ctx.predict(T, Y, best, t_est)
        dt = time.perf_counter() - t0
        lml = -funs.min(1)
        out[name] = {"gps": int(G), "m": int(T.shape[1]), "m_est": int(t_est.size), "starts": int(S),
                     "seconds": dt, "fits_per_s": G / dt, "lml_grad_evals": res["evals"], "rounds": res["rounds"],
                     "max_rel_lml_gap_vs_reference": float(np.max((g["lml_opt"] - lml) / np.abs(g["lml_opt"]))),
                     "reference_cpu_fit_seconds": float(np.sum(g["fit_seconds"]))}
    return out


def moments_sample(ctx, m, G=8):
    """compute_lstsq_matrices for G GPs at m = m' (state, ddt, ddt covariance, sqrtW by Newton-Schulz): host-call wall
    time (includes the D2H of two m'^2 matrices per GP) and per-class device milliseconds."""
    T, Y, _, _ = workload(G, m, 1, seed=321)
    theta = np.tile(np.log([1.5, 0.05, 1e-2]), (G, 1))
    t_est = np.linspace(0, 1, m)
    ctx.lstsq_weights(T[:1], Y[:1], theta[:1], t_est, 1e-8)          # warm-up / allocation
    ctx.profile_enable(True)
    t0 = time.perf_counter()
    state, ddt, cov, w, st, wst, wit = ctx.lstsq_weights(T, Y, theta, t_est, 1e-8)
    dt = time.perf_counter() - t0
    prof = ctx.profile_get()
    ctx.profile_enable(False)
    A = cov[0] + 1e-8 * np.eye(m)
    x = np.random.default_rng(0).standard_normal(m)
    resid = float(np.abs(w[0] @ (A @ (w[0] @ x)) - x).max() / np.abs(x).max())     # sqrtW (C + eta I) sqrtW x = x
    T_blocks = (m + 127) // 128
    ns_flops = float(wit.max() + 1) * G * 2.0 * 128 ** 3 * T_blocks * (T_blocks ** 2 + T_blocks * (T_blocks + 1))
    return {"gps": G, "m": m, "m_est": m, "seconds": dt, "device_ms": {k: round(v[0], 3) for k, v in prof.items() if v[1]},
            "sqrtw_iterations": int(wit.max()), "sqrtw_status_ok": int((wst == 0).sum()),
            "sqrtw_identity_residual": resid,
            "sqrtw_dmma_issued_tflops": ns_flops / (prof["sqrtw"][0] * 1e-3) / 1e12 if prof["sqrtw"][0] > 0 else None}


def cpu_baseline(T, Y, theta, budget_s=20.0):
    """Oracle port (sklearn-driven, as the reference) timed on this box's host cores, bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gp_oracle as orc
    from threadpoolctl import threadpool_info, threadpool_limits

    threadpool_limits(limits=os.cpu_count())
    b = np.array([(1e-5, 1e5), (1e-5, 1e2), (1e-16, 1e2)])
    gp = orc.OracleGP(tuple(b[0]), tuple(b[1]), tuple(b[2]), 0)
    gp.gpr.optimizer = None
    gp.fit(T[0], Y[0])
    times, k = [], 0
    t_start = time.perf_counter()
    while (time.perf_counter() - t_start < budget_s and k < 4) or k < 2:
        t0 = time.perf_counter()
        gp.lml_grad(theta[k])
        times.append(time.perf_counter() - t0)
        k += 1
    per = min(times[1:]) if len(times) > 1 else times[0]
    cores = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    return {"value": 1.0 / per, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{k} single-pair LML+grad evaluations at m={T.shape[1]} via sklearn "
                      "GaussianProcessRegressor.log_marginal_likelihood (best of the non-first)",
            "host_cpus": os.cpu_count()}


if __name__ == "__main__":
    main()
