"""Scratch GPU probe: bench.py's prediction roofline block alone (TRSM / Schur / means / std / sqrtW)."""
import json
import sys

sys.path.insert(0, ".")
import bench
from gpbo_pkg import pkg

ctx = pkg.default_context(0)
ctx.dmma_peak(100000)
peak = max(ctx.dmma_peak(100000)[0] for _ in range(3))
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
m = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
print(json.dumps(bench.prediction_roofline(pkg, ctx, peak, G=G, m=m)))
