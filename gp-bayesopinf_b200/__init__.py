"""gp_bayesopinf_b200 -- B200-native (sm_100a) implementation of GP-BayesOpInf's step2_fitgps hot path.

The directory is named ``gp-bayesopinf_b200`` (not an importable identifier); import it through the
repo-root helper ``gpbo_pkg.py`` (``from gpbo_pkg import pkg``), which registers it in ``sys.modules``
as ``gp_bayesopinf_b200``.

Public surface (mirrors the reference):
  gpkernels.GP_RBFW                      -- drop-in for codebase/gpkernels.py::GP_RBFW (sklearn path)
  step2_fitgps.fit_gaussian_processes    -- batched drop-in for */step2_fitgps.py
  step3_posterior.posterior_grid         -- posterior means / precisions of step 3 for a grid of regularizers (N3)
  _lib.Context                           -- ctypes handle on libgpbo.so (C ABI in include/gpbo.h)
"""

from . import _lib, gpkernels, sharding, step2_fitgps, step3_posterior, workload  # noqa: F401
from ._lib import Context, GpboError, default_context  # noqa: F401
from .gpkernels import GP_MaternW, GP_RBFW  # noqa: F401
from .step2_fitgps import fit_gaussian_processes, fit_gaussian_processes_multi  # noqa: F401

__version__ = "0.1.0"
