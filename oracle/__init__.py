"""CPU oracle of the pygent simulation-and-planning hot path (TEST INFRASTRUCTURE, NOT PRODUCT).

`pygent_oracle.c` restates the reference's algorithm (citations in its header); this module builds it
with gcc and wraps it with ctypes/numpy.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this package -- the product
(`pygent_b200/`) never does, and fails loudly without its CUDA library instead.

Parity is PINNED: `tests/test_oracle_golden.py` checks this oracle against outputs of the unmodified
reference frozen in `tests/golden/*.npz` by `oracle/make_golden.py`.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pygent_oracle.c")
LIB = os.path.join(HERE, "_build", "libpygent_oracle.so")

NMAX, MMAX = 8, 2
_i32, _i64, _f64 = ctypes.c_int32, ctypes.c_int64, ctypes.c_double
_dp = ctypes.POINTER(ctypes.c_double)


class Problem(ctypes.Structure):
    _fields_ = [("sys", _i32), ("n", _i32), ("m", _i32), ("S", _i32), ("integrator", _i32), ("lin_mode", _i32),
                ("dt", _f64), ("qx", _f64 * NMAX), ("ru", _f64 * MMAX), ("re", _f64 * MMAX), ("rd", _f64),
                ("qf", _f64 * NMAX), ("term_lim", _f64 * NMAX), ("terminal_cost", _f64),
                ("lx", _f64 * NMAX), ("c0", _f64), ("lf", _f64 * NMAX), ("cf0", _f64)]


class ILQRParams(ctypes.Structure):
    _fields_ = [("max_iters", _i32), ("reg_type", _i32), ("parallel", _i32), ("constrained", _i32), ("n_alpha", _i32),
                ("finite_diff", _i32), ("tol_grad", _f64), ("tol_fun", _f64), ("zmin", _f64), ("mu_min", _f64),
                ("mu_max", _f64), ("mu_d0", _f64), ("mu0", _f64), ("alphas", _f64 * 16), ("lims", _f64 * MMAX)]


class ILQRResult(ctypes.Structure):
    _fields_ = [("iters", _i32), ("status", _i32), ("cost", _f64), ("cost0", _f64), ("mu", _f64), ("mu_d", _f64),
                ("alpha", _f64)]


def build(force=False):
    """gcc recipe of the oracle (plain IEEE arithmetic: no FMA contraction, no fast-math)."""
    if os.path.isfile(LIB) and not force and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = ["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


_dll = None


def dll():
    global _dll
    if _dll is None:
        _dll = ctypes.CDLL(build())
        _dll.po_system_id.argtypes = [ctypes.c_char_p]
        _dll.po_system_id.restype = ctypes.c_int
        _dll.po_ilqr_forward.restype = ctypes.c_double
        _dll.po_ilqr_backward.restype = _i32
        _dll.po_max_threads.restype = ctypes.c_int
    return _dll


def _arr(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def system_id(name):
    sid = dll().po_system_id(name.encode())
    if sid < 0:
        raise KeyError(name)
    return sid


def system_dims(name):
    n, m, ab = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    dll().po_system_dims(system_id(name), ctypes.byref(n), ctypes.byref(m), ctypes.byref(ab))
    return n.value, m.value, bool(ab.value)


def make_problem(system, dt, S=4, integrator=0, lin_mode=None, qx=(), ru=(), re=(), rd=0.0, qf=(), term_lim=(),
                 terminal_cost=0.0, lx=(), c0=0.0, lf=(), cf0=0.0):
    """Cost family of the reference's examples:
    c = c0 + sum (lx_i x_i + qx_i x_i^2) + sum (ru_j + re_j exp(-rd t)) u_j^2 ; cf = cf0 + sum (lf_i x_i + qf_i x_i^2)."""
    n, m, has_ab = system_dims(system)
    p = Problem()
    p.sys, p.n, p.m, p.S, p.integrator = system_id(system), n, m, S, integrator
    p.lin_mode = (0 if has_ab else 1) if lin_mode is None else lin_mode
    p.dt, p.rd, p.terminal_cost, p.c0, p.cf0 = dt, rd, terminal_cost, c0, cf0
    for dst, src in ((p.qx, qx), (p.ru, ru), (p.re, re), (p.qf, qf), (p.term_lim, term_lim), (p.lx, lx), (p.lf, lf)):
        for i, v in enumerate(src):
            dst[i] = v
    return p


def f(system, x, u):
    n, m, _ = system_dims(system)
    x, u = _arr(x), _arr(u)
    out = np.empty(n)
    dll().po_f(system_id(system), _p(x), _p(u), _p(out))
    return out


def AB(system, x, u, fd=False):
    n, m, _ = system_dims(system)
    x, u = _arr(x), _arr(u)
    A, B = np.empty((n, n)), np.empty((n, m))
    (dll().po_AB_fd if fd else dll().po_AB)(system_id(system), _p(x), _p(u), _p(A), _p(B))
    return A, B


def rollout(p, x0, U=None, T=None, t0=0.0, history=True, add_final_cost=False, nthreads=1):
    """x0 [B,n], U [B,T,m] -> dict(X [B,T+1,n], xN [B,n], cost [B], terminated [B])."""
    x0 = _arr(x0)
    B = x0.shape[0]
    if U is not None:
        U = _arr(U)
        T = U.shape[1]
    X = np.empty((B, T + 1, p.n)) if history else None
    xN, cost, term = np.empty((B, p.n)), np.empty(B), np.empty(B, dtype=np.uint8)
    dll().po_rollout(ctypes.byref(p), _i64(B), _i32(T), _f64(t0), _p(x0), _p(U), _p(X), _p(xN), _p(cost), _p(term),
                     _i32(int(add_final_cost)), _i32(nthreads))
    return {"X": X, "xN": xN, "cost": cost, "terminated": term}


def ilqr_params(max_iters=500, tol_grad=1e-4, tol_fun=1e-7, reg_type=2, parallel=False, constrained=False,
                lims=(), alphas=None, finite_diff=False):
    q = ILQRParams()
    q.max_iters, q.reg_type, q.parallel, q.constrained = max_iters, reg_type, int(parallel), int(constrained)
    q.finite_diff = int(finite_diff)  # cost expansion by central differences (helpers.py:41-128)
    q.tol_grad, q.tol_fun, q.zmin = tol_grad, tol_fun, 0.0
    q.mu_min, q.mu_max, q.mu_d0, q.mu0 = 1e-6, 1e10, 1.6, 1e-6  # ilqr.py:143-148
    alphas = 10 ** np.linspace(0, -3, 11) if alphas is None else np.asarray(alphas, dtype=float)  # ilqr.py:149
    q.n_alpha = len(alphas)
    for i, a in enumerate(alphas):
        q.alphas[i] = a
    for i, l in enumerate(lims):
        q.lims[i] = l
    return q


def ilqr_backward(p, q, xx, uu, mu):
    """xx [T+1,n], uu [T,m] -> (KK [T,m,n], kk [T,m], sumV1, sumV2, gnorm, success)."""
    xx, uu = _arr(xx), _arr(uu)
    T = uu.shape[0]
    KK, kk, sums = np.zeros((T, p.m, p.n)), np.zeros((T, p.m)), np.zeros(3)
    ok = dll().po_ilqr_backward(ctypes.byref(p), ctypes.byref(q), _i32(T), _f64(mu), _p(xx), _p(uu), _p(KK), _p(kk), _p(sums))
    return KK, kk, sums[0], sums[1], sums[2], bool(ok)


def ilqr_forward(p, q, alpha, x0, xx, uu, KK, kk):
    """-> (xx_new [T+1,n], uu_new [T,m], cost, diverged, terminated)."""
    x0, xx, uu, KK, kk = _arr(x0), _arr(xx), _arr(uu), _arr(KK), _arr(kk)
    T = uu.shape[0]
    xo, uo = np.empty((T + 1, p.n)), np.empty((T, p.m))
    div, term = _i32(0), _i32(0)
    cost = dll().po_ilqr_forward(ctypes.byref(p), ctypes.byref(q), _i32(T), _f64(alpha), _p(x0), _p(xx), _p(uu), _p(KK),
                                 _p(kk), _i32(KK.shape[0]), _p(xo), _p(uo), ctypes.byref(div), ctypes.byref(term))
    return xo, uo, cost, bool(div.value), bool(term.value)


def ilqr_solve(p, q, x0, T, nthreads=1, trace=False):
    """x0 [B,n] -> dict(xx [B,T+1,n], uu [B,T,m], KK, kk, cost, cost0, iters, status, mu, alpha[, trace])."""
    x0 = _arr(x0)
    B = x0.shape[0]
    xx, uu = np.empty((B, T + 1, p.n)), np.empty((B, T, p.m))
    KK, kk = np.empty((B, T, p.m, p.n)), np.empty((B, T, p.m))
    res = (ILQRResult * B)()
    out = {}
    if trace:
        tr = np.full((B, q.max_iters, 4), np.nan)
        for b in range(B):
            dll().po_ilqr_solve_one(ctypes.byref(p), ctypes.byref(q), _i32(T), _p(x0[b]), _p(xx[b]), _p(uu[b]), _p(KK[b]),
                                    _p(kk[b]), ctypes.byref(res[b]), _p(tr[b]))
        out["trace"] = tr
    else:
        dll().po_ilqr_solve(ctypes.byref(p), ctypes.byref(q), _i64(B), _i32(T), _p(x0), _p(xx), _p(uu), _p(KK), _p(kk),
                            res, _i32(nthreads))
    out.update(xx=xx, uu=uu, KK=KK, kk=kk,
               cost=np.array([r.cost for r in res]), cost0=np.array([r.cost0 for r in res]),
               iters=np.array([r.iters for r in res]), status=np.array([r.status for r in res]),
               mu=np.array([r.mu for r in res]), alpha=np.array([r.alpha for r in res]))
    return out


def max_threads():
    return dll().po_max_threads()
