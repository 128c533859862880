"""`pygent` as an alias of `pygent_b200`: with this directory on PYTHONPATH,

    PYTHONPATH=/path/to/repo:/path/to/repo/pygent_b200/shims python examples/ilqr/cart_pole.py

runs a script written against the reference -- `from pygent.environments import CartPole`, `from pygent.algorithms.ilqr import
iLQR`, `pygent.algorithms.nmpc.NMPC`, `pygent.algorithms.mbrl.MBRL`, `pygent.helpers`, `pygent.agents`, `pygent.data`,
`pygent.nn_models` -- on the B200 kernels without touching it.  Only the modules of the hot path exist;
`pygent.algorithms.ddpg` / `nfq` and the gym wrapper are out of scope and raise ImportError.  Plotting and animation
(`Algorithm.plot`, `.animation`, `Environment.plot`, `.animation`) are accepted and do nothing."""
import importlib
import sys

import pygent_b200 as _impl

__path__ = []  # no submodules of its own: everything below is an alias
for _name in ("environments", "agents", "helpers", "data", "nn_models", "algorithms", "algorithms.core", "algorithms.ilqr",
              "algorithms.nmpc", "algorithms.mbrl"):
    _mod = importlib.import_module("pygent_b200." + _name)
    sys.modules[__name__ + "." + _name] = _mod
    if "." not in _name:
        globals()[_name] = _mod
__version__ = getattr(_impl, "__version__", "b200")
