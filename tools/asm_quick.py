"""Scratch GPU probe: stand-alone assembly at the bench shapes, all kinds, symmetric and general kernels, plus the
write-only streaming rate of the same buffers (cudaMemset through torch) for context.  Prints one line per shape."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from gpbo_pkg import pkg

label = sys.argv[1] if len(sys.argv) > 1 else "run"
only = sys.argv[2] if len(sys.argv) > 2 else ""
ctx = pkg.default_context(0)
dev = torch.device("cuda", 0)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
stream = torch.cuda.current_stream(dev)
for n1, n2, B in ((8192, 8192, 8), (16384, 16384, 2), (3200, 200, 64)):
    if only and only != str(n1):
        continue
    t1 = torch.sort(torch.rand(B, n1, dtype=torch.float64, device=dev), dim=1).values.contiguous()
    t2 = torch.sort(torch.rand(B, n2, dtype=torch.float64, device=dev), dim=1).values.contiguous()
    th = torch.log(torch.tensor([[1.5, 0.05, 1e-2]], dtype=torch.float64, device=dev)).repeat(B, 1).contiguous()
    o = torch.empty((B, n1, n2), dtype=torch.float64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o.zero_()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        o.zero_()
    e1.record()
    torch.cuda.synchronize()
    fill = 5 * o.numel() * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9 / peak
    res = {False: [], True: []}
    for sym in (False, True):
        for kind in range(7):
            if (kind in (0, 1, 5, 6) or sym) and n1 != n2:
                continue
            b2 = t1 if sym else t2
            args = (kind, t1.data_ptr(), n1, n1, b2.data_ptr(), n2, n2, th.data_ptr(), B, o.data_ptr(), stream.cuda_stream)
            ctx.assemble_device(*args)
            ctx.assemble_device(*args)
            torch.cuda.synchronize()
            ctx.profile_enable(True)
            for _ in range(10):
                ctx.assemble_device(*args)
            pms, pn = ctx.profile_get()["assemble"]
            ctx.profile_enable(False)
            res[sym].append(round(B * n1 * n2 * 8 / (pms / max(pn, 1) * 1e-3) / 1e9 / peak, 3))
    print(label, (n1, n2, B), "memset", round(fill, 3), "general", res[False], "sym", res[True], flush=True)
    del o
