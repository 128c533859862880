"""Symbolic derivation of the mechanical benchmark systems (Lagrange equations of the 2nd kind).

This is the product's own derivation of the models the reference ships as pickled sympy matrices
(`pygent/modeling_scripts/{cart_pole,acrobot,cart_pole_double_parallel,cart_pole_double_serial,
cart_pole_triple}.py` -> `c_files/*.p`).  Physics and parameter values are the reference's (cited
per system below); the derivation is not: instead of `sp.solve` on the expanded Euler-Lagrange
equations we build the mass matrix M(q) and the bias vector h(q, dq) once,

        M(q) ddq + h(q, dq) = Q(dq, u),

and obtain either the force/torque-input model  ddq = M^-1 (Q - h)  or the partially
feedback-linearised model (`linearized=True`, input = acceleration of the actuated coordinate a):

        ddq_a = u,      M_rr ddq_r = Q_r - h_r - M_ra u          (r = remaining coordinates)

by an explicit adjugate/determinant solve.  The result is the state-space right-hand side
f(x, u) with x = [q, dq]; A = df/dx and B = df/du are taken symbolically from the same f, exactly
as the reference does (`modeling_scripts/cart_pole.py:119-124`).

`python -m pygent_b200.modeling.derive` writes the expressions to `expr/<name>.txt`
(sympy `srepr`), which is what the package loads at run time -- the analogue of the reference's
`load_existing()` (`modeling_scripts/cart_pole.py:142-182`).
"""
import sympy as sp
from sympy import cos, sin


def _solve_small(M, rhs):
    """Explicit M^-1 rhs for a 1x1..3x3 symmetric mass matrix (adjugate / determinant)."""
    n = M.shape[0]
    if n == 1:
        return sp.Matrix([rhs[0] / M[0, 0]])
    det = M.det(method="berkowitz")
    adj = M.adjugate()
    return (adj * rhs) / det


def state_space(T, V, q, dq, Q, u_force, actuated, linearized, params):
    """f(x, u) for generalized coordinates q, velocities dq (plain sympy Symbols).

    T, V        kinetic / potential energy as expressions in q, dq
    Q           list of generalized forces; contains the Symbol `u_force` at index `actuated`
    linearized  True: input is the acceleration of coordinate `actuated`
    params      [(symbol, value)] substituted at the end
    Returns (f, xs, u) with xs = x1..x2N, u = Symbol('u').
    """
    nq = len(q)
    # exact rational parameters: structural zeros cancel exactly instead of leaving 1e-17 residues
    exact = [(s, sp.Rational(repr(v))) for s, v in params]
    T, V, Q = sp.sympify(T).subs(exact), sp.sympify(V).subs(exact), [sp.sympify(e).subs(exact) for e in Q]
    qv, dqv = sp.Matrix(q), sp.Matrix(dq)
    p = sp.Matrix([sp.diff(T, d) for d in dq])          # generalized momenta dT/ddq
    M = p.jacobian(dqv)                                  # mass matrix
    h = p.jacobian(qv) * dqv - sp.Matrix([sp.diff(T - V, c) for c in q])
    Qv = sp.Matrix(Q)
    u = sp.Symbol("u")
    if linearized:
        rest = [i for i in range(nq) if i != actuated]
        Mrr = M.extract(rest, rest)
        Mra = M.extract(rest, [actuated])
        rhs = sp.Matrix([Qv[i] - h[i] for i in rest]) - Mra * u
        sol = _solve_small(Mrr, rhs)
        ddq = [None] * nq
        ddq[actuated] = u
        for j, i in enumerate(rest):
            ddq[i] = sol[j]
    else:
        rhs = (Qv - h).subs(u_force, u)
        ddq = list(_solve_small(M, rhs))
    xs = sp.symbols("x1:%d" % (2 * nq + 1))
    sub = dict(zip(list(q) + list(dq), xs))
    f = sp.Matrix(list(dq) + list(ddq)).subs(sub)
    f = f.applyfunc(lambda e: _tidy(e, heavy=(linearized or nq <= 2)))
    return f, xs, u


def _tidy(e, heavy):
    """Canonical compact form with double-precision coefficients."""
    e = sp.simplify(e) if heavy else sp.trigsimp(sp.together(sp.expand(e)))
    num, den = sp.fraction(sp.together(e))
    if den != 1:
        # normalise so that the denominator's constant-most coefficient is O(1)
        lead = sp.Poly(den, *sorted(den.atoms(sp.sin, sp.cos), key=str)).coeffs()[-1] if den.atoms(sp.sin, sp.cos) else den
        num, den = sp.expand(num / lead), sp.expand(den / lead)
        return sp.N(num, 17) / sp.N(den, 17)
    return sp.N(sp.expand(num), 17)


def _coords(nq):
    q = sp.symbols("q0:%d" % nq, real=True)
    dq = sp.symbols("dq0:%d" % nq, real=True)
    return q, dq


def _vel(p, q, dq):
    return p.jacobian(sp.Matrix(q)) * sp.Matrix(dq)


def cart_pole(linearized=True):
    """Cart with one pole. Reference: modeling_scripts/cart_pole.py:11-88 (parameters :30-31)."""
    m0, m1, J1, a1, g, d0, d1, F = sp.symbols("m0 m1 J1 a1 g d0 d1 F")
    params = [(m0, 3.34), (m1, 0.3583), (J1, 0.0379999), (a1, 0.43), (g, 9.81), (d0, 0.1), (d1, 0.006588)]
    q, dq = _coords(2)
    p0 = sp.Matrix([q[0], 0])
    p1 = sp.Matrix([q[0] - a1 * sin(q[1]), a1 * cos(q[1])])
    v0, v1 = _vel(p0, q, dq), _vel(p1, q, dq)
    T = m0 / 2 * v0.dot(v0) + (m1 * v1.dot(v1) + J1 * dq[1] ** 2) / 2
    V = m1 * g * p1[1]
    Q = [F - d0 * dq[0], -d1 * dq[1]]
    return state_space(T, V, q, dq, Q, F, 0, linearized, params)


def acrobot(linearized=True):
    """Two-link arm, torque at the second joint. Reference: modeling_scripts/acrobot.py:11-80."""
    m0, m1, J0, J1, l0, l1, g, d0, d1, tau = sp.symbols("m0 m1 J0 J1 l0 l1 g d0 d1 tau")
    params = [(m0, 0.3583), (m1, 0.3583), (J0, 0.0379999), (J1, 0.0379999), (l0, 0.5), (l1, 0.5), (g, 9.81),
              (d0, 0.006588), (d1, 0.006588)]
    q, dq = _coords(2)
    half = sp.Rational(1, 2)
    p0 = sp.Matrix([-half * l0 * sin(q[0]), half * l0 * cos(q[0])])
    p1 = 2 * p0 + sp.Matrix([-half * l1 * sin(q[0] + q[1]), half * l1 * cos(q[0] + q[1])])
    v0, v1 = _vel(p0, q, dq), _vel(p1, q, dq)
    T = (m0 * v0.dot(v0) + J0 * dq[0] ** 2) / 2 + (m1 * v1.dot(v1) + J1 * (dq[0] + dq[1]) ** 2) / 2
    V = m0 * g * p0[1] + m1 * g * p1[1]
    Q = [-d0 * dq[0], tau - d1 * dq[1]]
    return state_space(T, V, q, dq, Q, tau, 1, linearized, params)


def cart_pole_double_parallel(linearized=True):
    """Cart with two independent poles. Reference: modeling_scripts/cart_pole_double_parallel.py:11-85."""
    m0, m1, m2, J1, J2, l1, l2, g, d0, d1, d2, F = sp.symbols("m0 m1 m2 J1 J2 l1 l2 g d0 d1 d2 F")
    params = [(m0, 3.34), (m1, 0.3583), (m2, 0.5383), (J1, 0.0379999), (J2, 0.12596), (l1, 0.5), (l2, 0.75),
              (g, 9.81), (d0, 0.1), (d1, 0.006588), (d2, 0.006588)]
    q, dq = _coords(3)
    half = sp.Rational(1, 2)
    p0 = sp.Matrix([q[0], 0])
    p1 = sp.Matrix([q[0] - half * l1 * sin(q[1]), half * l1 * cos(q[1])])
    p2 = sp.Matrix([q[0] - half * l2 * sin(q[2]), half * l2 * cos(q[2])])
    v0, v1, v2 = _vel(p0, q, dq), _vel(p1, q, dq), _vel(p2, q, dq)
    T = m0 / 2 * v0.dot(v0) + (m1 * v1.dot(v1) + J1 * dq[1] ** 2) / 2 + (m2 * v2.dot(v2) + J2 * dq[2] ** 2) / 2
    V = m1 * g * p1[1] + m2 * g * p2[1]
    Q = [F - d0 * dq[0], -d1 * dq[1], -d2 * dq[2]]
    return state_space(T, V, q, dq, Q, F, 0, linearized, params)


def cart_pole_double_serial(linearized=True):
    """Cart with a double pendulum. Reference: modeling_scripts/cart_pole_double_serial.py:13-88,
    parameter values :161-163."""
    m0, m1, m2, J1, J2, a1, a2, l1, g, d0, d1, d2, F = sp.symbols("m0 m1 m2 J1 J2 a1 a2 l1 g d0 d1 d2 F")
    params = [(m0, 3.34), (m1, 0.876), (m2, 0.938), (J1, 0.013), (J2, 0.024), (a1, 0.215), (a2, 0.269),
              (l1, 0.323), (g, 9.81), (d0, 0.1), (d1, 0.215), (d2, 0.002)]
    q, dq = _coords(3)
    p0 = sp.Matrix([q[0], 0])
    p1 = sp.Matrix([q[0] - a1 * sin(q[1]), a1 * cos(q[1])])
    p2 = sp.Matrix([q[0] - l1 * sin(q[1]) - a2 * sin(q[2]), l1 * cos(q[1]) + a2 * cos(q[2])])
    v0, v1, v2 = _vel(p0, q, dq), _vel(p1, q, dq), _vel(p2, q, dq)
    T = m0 / 2 * v0.dot(v0) + (m1 * v1.dot(v1) + J1 * dq[1] ** 2) / 2 + (m2 * v2.dot(v2) + J2 * dq[2] ** 2) / 2
    V = m1 * g * p1[1] + m2 * g * p2[1]
    Q = [F - d0 * dq[0], -d1 * dq[1] + d2 * (dq[2] - dq[1]), -d2 * (dq[2] - dq[1])]
    return state_space(T, V, q, dq, Q, F, 0, linearized, params)


def cart_pole_triple(linearized=True):
    """Cart with a triple pendulum (n = 8). Reference: modeling_scripts/cart_pole_triple.py:13-80,
    test-bench parameter values :178-182."""
    (m0, m1, m2, m3, J1, J2, J3, a1, a2, a3, l1, l2, g, d0, d1, d2, d3, F) = sp.symbols(
        "m0 m1 m2 m3 J1 J2 J3 a1 a2 a3 l1 l2 g d0 d1 d2 d3 F")
    params = [(m0, 3.34), (m1, 0.8512), (m2, 0.8973), (m3, 0.5519), (J1, 0.01980194), (J2, 0.02105375),
              (J3, 0.01818537), (a1, 0.20001517), (a2, 0.26890449), (a3, 0.21666087), (l1, 0.32), (l2, 0.419),
              (g, 9.81), (d0, 0.1), (d1, 0.00715294), (d2, 1.9497e-06), (d3, 0.00164642)]
    q, dq = _coords(4)
    p0 = sp.Matrix([q[0], 0])
    p1 = sp.Matrix([q[0] - a1 * sin(q[1]), a1 * cos(q[1])])
    p2 = sp.Matrix([q[0] - l1 * sin(q[1]) - a2 * sin(q[2]), l1 * cos(q[1]) + a2 * cos(q[2])])
    p3 = sp.Matrix([q[0] - l1 * sin(q[1]) - l2 * sin(q[2]) - a3 * sin(q[3]),
                    l1 * cos(q[1]) + l2 * cos(q[2]) + a3 * cos(q[3])])
    vs = [_vel(p, q, dq) for p in (p0, p1, p2, p3)]
    T = (m0 / 2 * vs[0].dot(vs[0]) + (m1 * vs[1].dot(vs[1]) + J1 * dq[1] ** 2) / 2
         + (m2 * vs[2].dot(vs[2]) + J2 * dq[2] ** 2) / 2 + (m3 * vs[3].dot(vs[3]) + J3 * dq[3] ** 2) / 2)
    V = m1 * g * p1[1] + m2 * g * p2[1] + m3 * g * p3[1]
    Q = [F - d0 * dq[0], -d1 * dq[1] + d2 * (dq[2] - dq[1]), -d2 * (dq[2] - dq[1]) + d3 * (dq[3] - dq[2]),
         -d3 * (dq[3] - dq[2])]
    return state_space(T, V, q, dq, Q, F, 0, linearized, params)


DERIVATIONS = {
    "cart_pole": cart_pole,
    "acrobot": acrobot,
    "cart_pole_double_parallel": cart_pole_double_parallel,
    "cart_pole_double_serial": cart_pole_double_serial,
    "cart_pole_triple": cart_pole_triple,
}
