"""Symbolic system registry: name -> (f, A, B) sympy matrices in x1..xn, u1..um.

Plays the role of the reference's `load_existing()` functions
(`pygent/modeling_scripts/cart_pole.py:142-182` etc.): the mechanical models are read from the
pre-derived expression files in `expr/` (see `derive.py`), the small hand-written ODEs of
`Pendulum`, `AngularPendulum`, `Car` and `MarBot` (`pygent/environments.py:338-349, 404-416,
740-754, 835-861`) are stated here directly as sympy expressions.

The returned `SystemModel` is the single source for (i) the numpy callables `ode/A/B` of the
Python environment classes and (ii) the generated CUDA device functions (`codegen.py`).
"""
import functools
import json
import os

import sympy as sp

EXPR_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "expr")


class SystemModel:
    """f, A, B as sympy matrices; `has_AB` False means the reference env ships no analytic
    Jacobians and iLQR linearises by central differences (`algorithms/ilqr.py:523-527`)."""

    def __init__(self, name, f, xs, us, has_AB, A=None, B=None):
        self.name = name
        self.xs = tuple(xs)
        self.us = tuple(us)
        self.n = len(xs)
        self.m = len(us)
        self.f = sp.Matrix(f)
        self.has_AB = has_AB
        self.A = A if A is not None else self.f.jacobian(list(xs))
        self.B = B if B is not None else self.f.jacobian(list(us))
        self._np = None

    def numpy_callables(self):
        """(ode(t, x, u), A(x, u), B(x, u)) as numpy functions, like the reference's lambdify
        fallback (`modeling_scripts/cart_pole.py:173-181`)."""
        if self._np is None:
            import numpy as np
            args = list(self.xs) + list(self.us)
            f_l = sp.lambdify(args, list(self.f), modules="numpy")
            A_l = sp.lambdify(args, self.A, modules="numpy")
            B_l = sp.lambdify(args, self.B, modules="numpy")
            self._np = (lambda t, x, u: np.array(f_l(*x, *u), dtype=float),
                        lambda x, u: np.array(A_l(*x, *u), dtype=float),
                        lambda x, u: np.array(B_l(*x, *u), dtype=float))
        return self._np


def _symbols(n, m):
    xs = sp.symbols("x1:%d" % (n + 1))
    us = sp.symbols("u1:%d" % (m + 1))
    return xs, us


def _from_file(stem):
    with open(os.path.join(EXPR_DIR, stem + ".json")) as fh:
        d = json.load(fh)
    n, m = d["n"], d["m"]
    xs, us = _symbols(n, m)
    loc = {str(s): s for s in xs}
    loc["u"] = us[0]
    par = lambda s: sp.sympify(s, locals=loc)
    f = sp.Matrix([par(e) for e in d["f"]])
    A = sp.Matrix([[par(e) for e in row] for row in d["A"]])
    B = sp.Matrix([[par(e) for e in row] for row in d["B"]])
    return SystemModel(stem, f, xs, us, True, A, B)


def _pendulum():
    # environments.py:338-349
    (x1, x2), (u1,) = xs, us = _symbols(2, 1)
    g, b = sp.Float(9.81), sp.Float(0.02)
    return SystemModel("pendulum", [x2, u1 + g * sp.sin(x1) - b * x2], xs, us, False)


def _angular_pendulum():
    # environments.py:404-416
    (x1, x2, x3), (u1,) = xs, us = _symbols(3, 1)
    g, b = sp.Float(9.81), sp.Float(0.02)
    return SystemModel("angular_pendulum", [-x2 * x3, x1 * x3, u1 + g * x2 - b * x3], xs, us, False)


def _car():
    # environments.py:740-754 (from the iLQG paper); m = 2
    (x1, x2, x3, x4), (u1, u2) = xs, us = _symbols(4, 2)
    d, h = sp.Float(2.0), sp.Float(0.03)
    fr = h * x4
    b = fr * sp.cos(u1) + d - sp.sqrt(d ** 2 - fr ** 2 * sp.sin(u1) ** 2)
    f = [1 / h * sp.cos(x3) * b, 1 / h * sp.sin(x3) * b, 1 / h * sp.asin(sp.sin(u1) * fr / d), u2]
    return SystemModel("car", f, xs, us, False)


def _marbot():
    # environments.py:835-861: M(q) ddq = F - C dq - G, solved in closed form (2x2)
    (x1, x2, x3, x4), (u1,) = xs, us = _symbols(4, 1)
    b, c, r, m_w, m_p, g = 0.2, 0.3, 0.05, 0.5, 3.0, 9.81
    J_w = 0.5 * m_w * r ** 2
    J_p = 1 / 12 * m_p * (b ** 2 + c ** 2)
    M = sp.Matrix([[2 * m_w + 2 * J_w / (r ** 2) + m_p, m_p * c * sp.cos(x2)],
                   [m_p * c * sp.cos(x2), J_p + m_p * c ** 2]])
    C = sp.Matrix([[0, -m_p * c * sp.sin(x2) * x4], [0, 0]])
    G = sp.Matrix([0, -m_p * c * g * sp.sin(x2)])
    F = sp.Matrix([2 / r * u1, 0])
    rhs = F - C * sp.Matrix([x3, x4]) - G
    det = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    ddq = sp.Matrix([[M[1, 1], -M[0, 1]], [-M[1, 0], M[0, 0]]]) * rhs / det
    return SystemModel("marbot", [x3, x4, ddq[0], ddq[1]], xs, us, False)


_HAND = {"pendulum": _pendulum, "angular_pendulum": _angular_pendulum, "car": _car, "marbot": _marbot}


@functools.lru_cache(maxsize=None)
def get_system(name):
    if name in _HAND:
        return _HAND[name]()
    if os.path.isfile(os.path.join(EXPR_DIR, name + ".json")):
        return _from_file(name)
    raise KeyError("unknown system %r" % name)


def available():
    files = sorted(f[:-5] for f in os.listdir(EXPR_DIR) if f.endswith(".json")) if os.path.isdir(EXPR_DIR) else []
    return sorted(_HAND) + files
