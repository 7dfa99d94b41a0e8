"""Import helper: registers the hyphenated package directory ``gp-bayesopinf_b200/`` as the module
``gp_bayesopinf_b200`` (used by tests, bench.py and __graft_entry__.py)."""

import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "gp-bayesopinf_b200")
NAME = "gp_bayesopinf_b200"


def _load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(
        NAME, os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    spec.loader.exec_module(mod)
    return mod


pkg = _load()
