"""Scratch GPU probe: bench.fit_reference_configs alone (full multi-start fits of the reference's five experiment
configurations from tests/golden), for A/B runs of library builds (GPBO_LIB)."""
import json
import os
import sys

sys.path.insert(0, ".")
import bench
from gpbo_pkg import pkg

ctx = pkg.default_context(0)
out = bench.fit_reference_configs(ctx)
print(os.environ.get("GPBO_LIB", "current"), json.dumps({k: {"seconds": round(v["seconds"], 5), "evals": v["lml_grad_evals"],
                                                              "chain": v["longest_chain"],
                                                              "gap": v["max_rel_lml_gap_vs_reference"]} for k, v in out.items()}))
