"""Import the UNMODIFIED reference (mpritzkoleit/pygent, /root/reference) in this container.

TEST / BASELINE TOOLING ONLY.  This module exists so that `oracle/make_golden.py` can run the real
reference and freeze its outputs under `tests/golden/`, and so that `bench.py`'s CPU legs (`cpu_baseline`,
`--impl reference`) can time the reference itself (oracle/ref_bench.py).  It is never imported by the product
(`pygent_b200/`), by `-m gpu` tests or by `smoke()`.  `/root/reference` does not exist on the GPU box; the
pip-installed copy oracle/_ref/ (oracle/install_ref.py; git-ignored, travels with the tree) does.

What it does:
  1. registers permissive stub modules for the reference's absent, import-time-only
     dependencies (gym, matplotlib.*, mpl_toolkits.*, cvxopt) -- `environments.py:4-9`,
     `algorithms/ilqr.py:3,15-16`;
  2. swaps the `pickle` attribute of the five `pygent.modeling_scripts.*` modules for a shim
     Unpickler that can read the old-sympy `c_files/*.p` model pickles under sympy 1.14;
  3. offers `rk4_step(S)`: a replacement for `StateSpaceModel.step` that is identical to
     `environments.py:253-276` except that the `solve_ivp` line becomes S classical RK4 substeps.
"""
import contextlib
import io
import os
import pickle
import sys
import types

_INSTALLED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_root():
    """Directory that contains the reference's `pygent` package: $PYGENT_REFERENCE, /root/reference, or oracle/_ref."""
    for cand in (os.environ.get("PYGENT_REFERENCE"), "/root/reference", _INSTALLED):
        if cand and os.path.isfile(os.path.join(cand, "pygent", "algorithms", "ilqr.py")):
            return cand
    return None


REFERENCE_ROOT = reference_root() or "/root/reference"


class _Stub(types.ModuleType):
    """Module stub: any attribute is another stub, calling it returns itself."""

    def __init__(self, name):
        super().__init__(name)
        self.__path__ = []
        self.options = {}

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        child = _Stub(self.__name__ + "." + item)
        setattr(self, item, child)
        return child

    def __call__(self, *a, **k):
        return self


_STUBS = ["gym", "matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.patches",
          "matplotlib.ticker", "matplotlib.colors", "mpl_toolkits", "mpl_toolkits.mplot3d", "cvxopt",
          "cvxopt.solvers"]


class _MatrixHolder:
    def __setstate__(self, state):
        self.__dict__.update(state)


class _ShimUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        import sympy
        from sympy.core.singleton import Singleton
        if module.startswith("sympy.matrices") and name in ("MutableDenseMatrix", "ImmutableDenseMatrix"):
            return _MatrixHolder
        cls = super().find_class(module, name)
        if isinstance(cls, Singleton):
            return lambda *a, **k: cls()
        return cls


def shim_load(fh):
    import sympy as sp
    obj = _ShimUnpickler(fh).load()
    if isinstance(obj, _MatrixHolder):
        return sp.Matrix(obj.rows, obj.cols, list(obj._mat))
    return obj


_loaded = None


def load_reference():
    """Returns the imported `pygent` package of the reference (cached)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError("reference tree %s not present (run `python -m oracle.install_ref` where /root/reference exists)" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = _Stub(name)
    sys.modules["cvxopt"].solvers = sys.modules["cvxopt.solvers"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with contextlib.redirect_stdout(io.StringIO()):
        import pygent.modeling_scripts.cart_pole as m1
        import pygent.modeling_scripts.acrobot as m2
        import pygent.modeling_scripts.cart_pole_double_parallel as m3
        import pygent.modeling_scripts.cart_pole_double_serial as m4
        import pygent.modeling_scripts.cart_pole_triple as m5
        for m in (m1, m2, m3, m4, m5):
            m.pickle = types.SimpleNamespace(load=shim_load, dump=pickle.dump)
        import pygent
        import pygent.environments
        import pygent.algorithms.ilqr
    _loaded = pygent
    return pygent


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def load_pickled_model(stem):
    """sympy (ode, A, B) of `c_files/<stem>_{ode,A,B}.p` as shipped (parameters not substituted)."""
    out = []
    for part in ("ode", "A", "B"):
        p = os.path.join(REFERENCE_ROOT, "pygent", "modeling_scripts", "c_files", "%s_%s.p" % (stem, part))
        with open(p, "rb") as fh:
            out.append(shim_load(fh))
    return tuple(out)


def rk4_step(substeps):
    """Fixed-step replacement for `StateSpaceModel.step` (environments.py:242-276).

    Everything but the integrator call is the reference's own statement sequence; the
    `solve_ivp(..., 'RK45')` line is replaced by `substeps` classical RK4 steps of dt/substeps.
    """
    import numpy as np

    def step(self, *args):
        self.x_ = self.x
        self.o_ = self.o
        if args.__len__() == 2:
            u = args[0]
            dt = args[1]
        elif args.__len__() == 1:
            u = args[0]
            dt = self.dt
        h = dt / substeps
        x = np.array(self.x_, dtype=float)
        for _ in range(substeps):
            k1 = np.asarray(self.ode(0, x, u), dtype=float)
            k2 = np.asarray(self.ode(0, x + 0.5 * h * k1, u), dtype=float)
            k3 = np.asarray(self.ode(0, x + 0.5 * h * k2, u), dtype=float)
            k4 = np.asarray(self.ode(0, x + h * k3, u), dtype=float)
            x = x + (h / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
        self.x = list(x)
        self.o = self.observe(self.x)
        self.history = np.concatenate((self.history, np.array([self.x])))
        self.tt.extend([self.tt[-1] + dt])
        self.terminated = self.terminate(self.x)
        t = self.tt[-1]
        c = (self.cost(self.x_, u, self.x, t, np) + self.terminal_cost * self.terminated) * dt
        return c

    return step


@contextlib.contextmanager
def fixed_step(substeps):
    """Context manager: reference runs with RK4 x `substeps` instead of adaptive RK45."""
    pygent = load_reference()
    cls = pygent.environments.StateSpaceModel
    orig = cls.step
    cls.step = rk4_step(substeps)
    try:
        yield
    finally:
        cls.step = orig


def install_qp_standin():
    """cvxopt is absent from this image, and `constrained=True` -- the branch every shipped examples/ilqr/*.py uses -- solves
    its box QP with it (ilqr.py:303-316).  This installs a stand-in module with exactly the three names the reference touches:
    `matrix`, `solvers.options` and `solvers.qp(P, q, G, h)` = argmin 1/2 x'Px + q'x subject to G x <= h, returned as
    {'x': column}.  The reference only ever passes the box pattern G = [I; -I], h = [up; lo]; the stand-in checks that and
    solves the bounded least-squares form of the problem with scipy's BVLS (an active-set method: exact to rounding, where
    cvxopt's interior point stops ~1e-8 inside the bounds).  Everything around the QP -- regularisation, the clamp detection
    with its lower-bound sign quirk, the zeroed feedback rows -- is the reference's own code, so goldens made with this pin the
    constrained branch up to the QP solver's tolerance.  Generator-side only."""
    import numpy as np
    from scipy.optimize import lsq_linear

    def qp(P, q, G, h, *a, **k):
        P, q = np.atleast_2d(np.asarray(P, dtype=float)), np.asarray(q, dtype=float).reshape(-1)
        G, h = np.atleast_2d(np.asarray(G, dtype=float)), np.asarray(h, dtype=float).reshape(-1)
        m = len(q)
        if G.shape != (2 * m, m) or not (np.array_equal(G[:m], np.eye(m)) and np.array_equal(G[m:], -np.eye(m))):
            raise NotImplementedError("qp stand-in: box constraints only (the reference's G = [I; -I])")
        L = np.linalg.cholesky(P)                      # 1/2 x'Px + q'x = 1/2 |L'x + L^-1 q|^2 + const
        r = lsq_linear(L.T, -np.linalg.solve(L, q), bounds=(-h[m:], h[:m]), method="bvls", tol=1e-15, max_iter=200)
        return {"x": r.x.reshape(m, 1), "status": "optimal"}
    mod, solvers = types.ModuleType("cvxopt"), types.ModuleType("cvxopt.solvers")
    mod.matrix = lambda a, *x, **k: np.array(a, dtype=float)
    solvers.options, solvers.qp = {}, qp
    mod.solvers = solvers
    sys.modules["cvxopt"], sys.modules["cvxopt.solvers"] = mod, solvers
    for name, m_ in list(sys.modules.items()):  # modules of the reference that were imported with the permissive stub
        if name.startswith("pygent") and getattr(m_, "opt", None) is not None and hasattr(m_, "iLQR"):
            m_.opt = mod
    return mod

