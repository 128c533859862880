"""Turn user callables (costs, ODEs) into sympy expressions.

The reference obtains the cost expansion by calling the user's cost with sympy symbols and `mod=sp`
(`pygent/algorithms/ilqr.py:539-555`).  We do the same to obtain the expressions the CUDA code is
generated from.  Callables that use numpy ufuncs directly (`np.sin(x1)`, as `Pendulum.ode` does,
`pygent/environments.py:338-349`) are traced with `Sym` proxies, which answer numpy's ufunc
protocol with the sympy function of the same name.
"""
import inspect

import numpy as np
import sympy as sp

_UFUNCS = {"sin": sp.sin, "cos": sp.cos, "tan": sp.tan, "exp": sp.exp, "log": sp.log, "sqrt": sp.sqrt,
           "arcsin": sp.asin, "arccos": sp.acos, "arctan": sp.atan, "absolute": sp.Abs, "fabs": sp.Abs,
           "tanh": sp.tanh, "sinh": sp.sinh, "cosh": sp.cosh, "sign": sp.sign, "square": lambda a: a ** 2,
           "negative": lambda a: -a, "positive": lambda a: a, "arctan2": sp.atan2,
           "add": lambda a, b: a + b, "subtract": lambda a, b: a - b, "multiply": lambda a, b: a * b,
           "true_divide": lambda a, b: a / b, "divide": lambda a, b: a / b, "power": lambda a, b: a ** b}


def _unwrap(v):
    if isinstance(v, Sym):
        return v.e
    if isinstance(v, (np.floating, np.integer)):
        return sp.sympify(v.item())
    return sp.sympify(v)


class Sym:
    """Arithmetic proxy around a sympy expression that numpy ufuncs can be applied to."""
    __array_priority__ = 1000

    def __init__(self, e):
        self.e = sp.sympify(e)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        fn = _UFUNCS.get(ufunc.__name__)
        if method != "__call__" or fn is None:
            return NotImplemented
        return Sym(fn(*[_unwrap(i) for i in inputs]))

    def _bin(op):
        def f(self, other):
            return Sym(op(self.e, _unwrap(other)))
        return f

    def _rbin(op):
        def f(self, other):
            return Sym(op(_unwrap(other), self.e))
        return f

    __add__, __radd__ = _bin(lambda a, b: a + b), _rbin(lambda a, b: a + b)
    __sub__, __rsub__ = _bin(lambda a, b: a - b), _rbin(lambda a, b: a - b)
    __mul__, __rmul__ = _bin(lambda a, b: a * b), _rbin(lambda a, b: a * b)
    __truediv__, __rtruediv__ = _bin(lambda a, b: a / b), _rbin(lambda a, b: a / b)
    __pow__, __rpow__ = _bin(lambda a, b: a ** b), _rbin(lambda a, b: a ** b)

    def __neg__(self):
        return Sym(-self.e)

    def __pos__(self):
        return self

    def __abs__(self):
        return Sym(sp.Abs(self.e))



# numpy falls back to calling a method of the ufunc's name on unknown objects (object arrays)
for _name, _fn in (("sin", sp.sin), ("cos", sp.cos), ("tan", sp.tan), ("exp", sp.exp), ("log", sp.log),
                   ("sqrt", sp.sqrt), ("arcsin", sp.asin), ("arccos", sp.acos), ("arctan", sp.atan), ("tanh", sp.tanh)):
    setattr(Sym, _name, (lambda fn: lambda self: Sym(fn(self.e)))(_fn))
del _name, _fn


def to_expr(v):
    """Result of a traced call -> sympy expression (scalars) or list of expressions."""
    if isinstance(v, (list, tuple, np.ndarray)):
        return [to_expr(e) for e in (v.tolist() if isinstance(v, np.ndarray) else v)]
    if isinstance(v, sp.MatrixBase):
        return [to_expr(e) for e in v]
    return _unwrap(v)


def trace(fn, *args):
    """Call fn with sympy symbols; on failure (numpy ufunc on a Symbol) retry with Sym proxies."""
    try:
        return to_expr(fn(*args))
    except (TypeError, AttributeError, ValueError, sp.SympifyError):
        wrap = lambda a: [Sym(s) for s in a] if isinstance(a, (list, tuple)) else (Sym(a) if isinstance(a, sp.Basic) else a)
        return to_expr(fn(*[wrap(a) for a in args]))


def bind_cost(cost):
    """The reference's cost-signature dispatch (`pygent/environments.py:206-235`): a user cost of 1-5
    arguments becomes cost(x_, u_, x, t, mod), decided by argument count and the names `mod` / `t`."""
    params = inspect.signature(cost).parameters
    k = len(params)
    if k == 1:
        return lambda x_, u_, x, t, mod: cost(x_)
    if k == 2:
        if "mod" in params:
            return lambda x_, u_, x, t, mod: cost(x_, mod)
        if "t" in params:
            return lambda x_, u_, x, t, mod: cost(x_, t)
        return lambda x_, u_, x, t, mod: cost(x_, u_)
    if k == 3:
        if "mod" in params:
            return lambda x_, u_, x, t, mod: cost(x_, u_, mod)
        if "t" in params:
            return lambda x_, u_, x, t, mod: cost(x_, u_, t)
        return lambda x_, u_, x, t, mod: cost(x_, u_, x)
    if k == 4:
        if "mod" in params and "t" in params:
            return lambda x_, u_, x, t, mod: cost(x_, u_, t, mod)
        if "mod" in params:
            return lambda x_, u_, x, t, mod: cost(x_, u_, x, mod)
        return lambda x_, u_, x, t, mod: cost(x_, u_, x, t)
    if k == 5:
        return cost
    raise TypeError("Cost function must be of the form c(x_, u_, x, t, mod), where mod is numpy/sympy.")


def bind_fcost(fcost):
    """`iLQR.__init__` final-cost dispatch (`pygent/algorithms/ilqr.py:119-125`)."""
    k = len(inspect.signature(fcost).parameters)
    if k == 1:
        return lambda x, mod: fcost(x)
    if k == 2:
        return fcost
    raise TypeError("final cost must be fcost(x) or fcost(x, mod)")
