"""Scratch GPU probe (not the contract bench): DMMA peak, cuBLAS DGEMM, and per-kernel-class timing of one
batched LML+grad evaluation."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
from gpbo_pkg import pkg
import gp_oracle as orc

m = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148
ctx = pkg.default_context(0)
out = {}
SKIP_PEAKS = len(sys.argv) > 3
for it in (() if SKIP_PEAKS else (20000, 200000)):
    tf, ms = ctx.dmma_peak(it)
    out[f"dmma_peak_tflops_{it}"] = (tf, ms)
# cuBLAS DGEMM as the 'achievable library' line
n = 8192 if not SKIP_PEAKS else 256
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
torch.matmul(a, b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    torch.matmul(a, b)
e1.record()
torch.cuda.synchronize()
out["cublas_dgemm_tflops"] = 3 * 2 * n**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
del a, b
torch.cuda.empty_cache()

r = 4
t, y = orc.synthetic_trajectories(r, m, seed=0)
T = np.tile(t, (r, 1))
rng = np.random.default_rng(0)
theta = np.log(np.array([1.0, 0.05, 1e-2]))[None, :] + 0.2 * rng.standard_normal((B, 3))
gp_of = (np.arange(B) % r).astype(np.int32)
dev = torch.device("cuda", 0)
Td, Yd, thd = (torch.as_tensor(x, device=dev) for x in (T, y, theta))
gpd = torch.as_tensor(gp_of, device=dev)
lml = torch.empty(B, dtype=torch.float64, device=dev)
grad = torch.empty(B, 3, dtype=torch.float64, device=dev)
st = torch.empty(B, dtype=torch.int32, device=dev)
torch.cuda.synchronize()
args = (Td.data_ptr(), Yd.data_ptr(), r, m, thd.data_ptr(), gpd.data_ptr(), B, lml.data_ptr(), grad.data_ptr(), st.data_ptr(), 0)
ctx.lml_grad_device(*args)  # warm-up
ctx.profile_enable(True)
t0 = time.perf_counter()
ctx.lml_grad_device(*args)
dt = time.perf_counter() - t0
prof = ctx.profile_get()
ctx.profile_enable(False)
out["m"], out["B"], out["wave_capacity"] = m, B, ctx.wave_capacity(m)
out["eval_seconds"] = dt
out["evals_per_s"] = B / dt
out["tflops_m3"] = B * float(m) ** 3 / dt / 1e12
out["profile_ms"] = prof
out["status_bad"] = int((st != 0).sum().item())
l0, g0, _ = orc.np_lml_grad(t, y[gp_of[0]], theta[0]) if m <= 2048 else (None, None, None)
out["lml0"] = (float(lml[0].item()), l0)
print(json.dumps(out))
