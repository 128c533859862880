"""sympy -> CUDA device functions: the product's counterpart of the reference's native-codegen seam.

The reference hands sympy matrices to the third-party `sympy_to_c.convert_to_c(args, expr,
cfilepath=...)` to obtain compiled callables for the ODE, its Jacobians and the cost expansion
(`pygent/modeling_scripts/cart_pole.py:159-170`, `pygent/algorithms/ilqr.py:558-571`).  Here the same
expressions become ONE generated header, `struct Problem`, with templated `__device__` functions

    f(x,u,dx)  AB(x,u,A,B)  cost(x,u,xn,t)  fcost(x)  cost_derivs(...)  fcost_derivs(...)  terminate(x)

plus structural non-zero masks, against which the hand-written kernels in `csrc/pg_kernels.cuh` are
instantiated (one shared library per problem, see `build.py`).

Code generation: every sin/cos pair of the same argument is emitted as one `pg::sincos`, the
remaining algebra goes through `sympy.cse`; numeric literals are wrapped in `T(...)` so that the
float instantiation contains no double arithmetic.
"""
import hashlib

import sympy as sp
from sympy.printing.c import C99CodePrinter

ABI_VERSION = 5
CODEGEN_VERSION = 5  # bump to invalidate cached libraries when the emitted code changes

_FUNCS = {"sin": "pg::sin", "cos": "pg::cos", "tan": "pg::tan", "exp": "pg::exp", "log": "pg::log",
          "sqrt": "pg::sqrt", "asin": "pg::asin", "acos": "pg::acos", "atan": "pg::atan",
          "Abs": "pg::abs", "tanh": "pg::tanh", "sinh": "pg::sinh", "cosh": "pg::cosh", "sign": "pg::sign",
          "atan2": "pg::atan2"}


class _Printer(C99CodePrinter):
    def __init__(self):
        super().__init__({"user_functions": dict(_FUNCS)})

    def _print_Float(self, e):
        return "T(%s)" % repr(float(e))

    def _print_Integer(self, e):
        return "T(%d)" % int(e)

    def _print_Rational(self, e):
        return "T(%d.0/%d.0)" % (e.p, e.q)

    def _print_NumberSymbol(self, e):
        return "T(%s)" % repr(float(e))

    _print_Pi = _print_Exp1 = _print_NumberSymbol

    def _print_Pow(self, e):
        b, x = e.base, e.exp
        if x.is_Integer:
            k = int(x)
            if k == -1:
                return "(T(1)/(%s))" % self._print(b)
            if k < 0:
                return "(T(1)/pg::powi<%d>(%s))" % (-k, self._print(b))
            return "pg::powi<%d>(%s)" % (k, self._print(b))
        if x == sp.Rational(1, 2):
            return "pg::sqrt(%s)" % self._print(b)
        if x == sp.Rational(-1, 2):
            return "(T(1)/pg::sqrt(%s))" % self._print(b)
        return "pg::pow(%s, %s)" % (self._print(b), self._print(x))


_printer = _Printer()


def _cxx(e):
    return _printer.doprint(e)


def _pair_trig(exprs, prefix):
    """Replace sin(arg)/cos(arg) by symbols; returns (exprs, [(arg, sin_sym|None, cos_sym|None)]).

    Innermost arguments first, so that nested trig (rare) is emitted in dependency order."""
    defs = []
    exprs = [sp.sympify(e) for e in exprs]
    k = 0
    while True:
        atoms = set()
        for e in exprs:
            atoms |= e.atoms(sp.sin, sp.cos)
        for a, _, _ in defs:
            atoms |= a.atoms(sp.sin, sp.cos)
        leaf = [a for a in atoms if not a.args[0].atoms(sp.sin, sp.cos)]
        if not leaf:
            break
        by_arg = {}
        for a in leaf:
            by_arg.setdefault(a.args[0], set()).add(type(a))
        rep = {}
        for arg in sorted(by_arg, key=sp.default_sort_key):
            s = sp.Symbol("%ssn%d" % (prefix, k)) if sp.sin in by_arg[arg] else None
            c = sp.Symbol("%scs%d" % (prefix, k)) if sp.cos in by_arg[arg] else None
            k += 1
            if s is not None:
                rep[sp.sin(arg)] = s
            if c is not None:
                rep[sp.cos(arg)] = c
            defs.append((arg, s, c))
        exprs = [e.xreplace(rep) for e in exprs]
        # trig args that themselves contain already-extracted trig
        defs = [(a.xreplace(rep), s, c) for a, s, c in defs]
    return exprs, defs


def _needs_register(fl):
    """A double literal whose low 32 bits are non-zero cannot be an instruction immediate."""
    import struct
    return struct.unpack("<Q", struct.pack("<d", float(fl)))[0] & 0xFFFFFFFF != 0


def hoistable_constants(exprs, limit=8):
    """Float literals of the hot expressions that are worth keeping in registers (Ctx::kN)."""
    found = []
    for e in exprs:
        for fl in sorted(sp.sympify(e).atoms(sp.Float), key=lambda v: float(v)):
            if _needs_register(fl) and all(float(fl) != float(o) for o in found):
                found.append(fl)
    return found[:limit]


def anchored_trig_args(f_exprs, xs):
    """The sin/cos arguments of the ODE right-hand side if ALL of them are affine in the state with numeric
    coefficients (x2, x1 + x2, ...): [(arg, needs_sin, needs_cos)], in the order `emit_body` numbers them;
    else [].  For such systems the integrator keeps sin/cos of each argument as an "anchor" and obtains the
    values at the RK stages by rotating the anchor through the small argument increment (pg_math.cuh)."""
    xs = list(xs)
    _, trig = _pair_trig([sp.sympify(e) for e in f_exprs], "w")
    out = []
    for arg, s, c in trig:
        if not arg.free_symbols or not arg.free_symbols <= set(xs):
            return []
        if arg.atoms(sp.Function) or not all(sp.diff(arg, x).is_number for x in xs):
            return []
        out.append((arg, s is not None, c is not None))
    return out


def emit_body(outputs, sym_names, prefix="w", hoisted=(), given_trig=False):
    """C++ statements computing `outputs` = [(lhs_string, sympy expr)].

    sym_names: {sympy Symbol: C++ lvalue string}.  Zero expressions are still assigned (callers
    rely on every output being written).  `hoisted` float literals are read from the context
    (`m.kN`) instead of being emitted as immediates."""
    ren = {s: sp.Symbol(n) for s, n in sym_names.items()}
    exprs = [sp.sympify(e).xreplace(ren) for _, e in outputs]
    if hoisted:
        hmap = {}
        for e in exprs:
            for fl in e.atoms(sp.Float):
                for i, h in enumerate(hoisted):
                    if float(fl) == float(h):
                        hmap[fl] = sp.Symbol("m.k%d" % i)
        exprs = [e.xreplace(hmap) for e in exprs]
    exprs, trig = _pair_trig(exprs, prefix)
    lines = []
    # trig arguments may share sub-expressions with the algebra; keep it simple: emit them first
    for j, (arg, s, c) in enumerate(trig):
        if given_trig:  # sin/cos of the j-th argument are inputs (sn[j], cs[j])
            if s is not None:
                lines.append("const T %s = sn[%d];" % (s, j))
            if c is not None:
                lines.append("const T %s = cs[%d];" % (c, j))
        elif s is not None and c is not None:
            lines.append("T %s, %s; m.sincos(%s, %s, %s);" % (s, c, _cxx(arg), s, c))
        elif s is not None:
            lines.append("const T %s = m.sin(%s);" % (s, _cxx(arg)))
        else:
            lines.append("const T %s = m.cos(%s);" % (c, _cxx(arg)))
    repl, red = sp.cse(exprs, symbols=sp.numbered_symbols(prefix + "t"), optimizations="basic", order="none")
    for s, e in repl:
        lines.append("const T %s = %s;" % (s, _cxx(e)))
    for (lhs, _), e in zip(outputs, red):
        lines.append("%s = %s;" % (lhs, _cxx(e)))
    return lines


def _mask_fn(name, mat):
    """Structural non-zero mask as a constexpr function (usable with loop indices in device code)."""
    mat = list(mat)
    body = "{" + ", ".join("1" if e != 0 else "0" for e in mat) + "}"
    return ("    __host__ __device__ static constexpr bool %s(int i) { constexpr unsigned char m[%d] = %s; return m[i] != 0; }"
            % (name, len(mat), body))


def _fn(sig, body_lines, indent="    "):
    name, rest = sig.split("(", 1)
    sig = name + "(const Ctx<T>& m, " + rest
    out = [indent + "template <typename T> __device__ __forceinline__ static " + sig + " {", indent + "    (void)m;"]
    out += [indent + "    " + l for l in body_lines]
    out.append(indent + "}")
    return out


class ProblemSpec:
    """System dynamics + running cost + final cost + termination box = one compiled problem.

    cost   sympy expr in xs, us, xn (next state), t    (the reference's `cost(x_, u_, x, t, mod)`,
           `environments.py:206-235`; NOT multiplied by dt)
    fcost  sympy expr in xs                            (`ilqr.py:119-125`)
    term   [(state index, limit)] : terminated = any(|x_i| > limit)   (`environments.py:477-482` ...)
    """

    def __init__(self, system, cost=None, fcost=None, term=(), x_is_angle=None):
        self.system = system
        n, m = system.n, system.m
        self.xs, self.us = system.xs, system.us
        self.xn = sp.symbols("xn1:%d" % (n + 1))
        self.t = sp.Symbol("t")
        self.cost = sp.sympify(cost if cost is not None else 0)
        self.fcost = sp.sympify(fcost if fcost is not None else 0)
        self.term = tuple((int(i), float(l)) for i, l in term)
        self.x_is_angle = tuple(bool(a) for a in (x_is_angle if x_is_angle is not None else [False] * n))
        extra = (self.cost.free_symbols | self.fcost.free_symbols) - set(self.xs) - set(self.us) - set(self.xn) - {self.t}
        if extra:
            raise ValueError("cost uses unknown symbols %s" % sorted(map(str, extra)))
        if self.fcost.free_symbols - set(self.xs):
            raise ValueError("final cost may depend on the state only")

    @property
    def cost_uses_xnext(self):
        return bool(self.cost.free_symbols & set(self.xn))

    def key(self, extra=""):
        h = hashlib.sha1()
        s = self.system
        parts = [str(ABI_VERSION), str(CODEGEN_VERSION), s.name, sp.srepr(s.f), sp.srepr(s.A), sp.srepr(s.B), str(s.has_AB),
                 sp.srepr(self.cost), sp.srepr(self.fcost), repr(self.term), repr(self.x_is_angle), extra]
        h.update("\n".join(parts).encode())
        return h.hexdigest()[:16]

    def header(self, key):
        s = self.system
        n, m = s.n, s.m
        names = {x: "x[%d]" % i for i, x in enumerate(self.xs)}
        names.update({u: "u[%d]" % j for j, u in enumerate(self.us)})
        names.update({x: "xn[%d]" % i for i, x in enumerate(self.xn)})
        names[self.t] = "t"
        xs, us = list(self.xs), list(self.us)
        c = sp.Matrix([[self.cost]])
        if self.cost_uses_xnext:
            # iLQR's expansion evaluates the running cost with x' = None (ilqr.py:183): no derivative exists
            cx = cu = Cxx = Cuu = Cxu = None
        else:
            cx = c.jacobian(xs)
            cu = c.jacobian(us)
            Cxx = cx.jacobian(xs)
            Cuu = cu.jacobian(us)
            Cxu = cx.jacobian(us)
        cf = sp.Matrix([[self.fcost]])
        cfx = cf.jacobian(xs)
        Cfxx = cfx.jacobian(xs)
        n_obs = n + sum(self.x_is_angle)

        L = ["// generated by pygent_b200.codegen -- do not edit", "#pragma once",
             '#define PG_PROBLEM_NAME "%s"' % s.name, '#define PG_PROBLEM_KEY "%s"' % key,
             "__device__ const double PG_PROBLEM_CONSTS[%d] = {%s};" % (max(1, len(hoistable_constants(list(s.f)))), ", ".join(repr(float(h)) for h in hoistable_constants(list(s.f))) or "0.0"),
             "struct Problem {",
             "    static constexpr int N = %d, M = %d, NOBS = %d;" % (n, m, n_obs),
             "    static constexpr bool HAS_AB = %s;" % ("true" if s.has_AB else "false"),
             "    static constexpr bool HAS_COST_DERIVS = %s;" % ("false" if cx is None else "true"),
             "    static constexpr bool COST_USES_XNEXT = %s;" % ("true" if self.cost_uses_xnext else "false"),
             _mask_fn("x_is_angle", [int(a) for a in self.x_is_angle]),
             _mask_fn("a_nz", s.A), _mask_fn("b_nz", s.B)]
        if cx is not None:
            L += [_mask_fn("cxx_nz", Cxx), _mask_fn("cuu_nz", Cuu), _mask_fn("cxu_nz", Cxu)]
        else:
            L += [_mask_fn("cxx_nz", [1] * (n * n)), _mask_fn("cuu_nz", [1] * (m * m)), _mask_fn("cxu_nz", [1] * (n * m))]
        L.append(_mask_fn("cfxx_nz", Cfxx))

        # per-thread context: sin/cos coefficients (pg::MathCtx) + the literals of f, all held in registers
        hoisted = hoistable_constants(list(s.f))
        L += ["    template <typename T> struct Ctx : pg::MathCtx<T> {"]
        if hoisted:
            L += ["        T %s;" % ", ".join("k%d" % i for i in range(len(hoisted)))]
        L += ["        __device__ __forceinline__ Ctx() {"]
        L += ["            k%d = pg::hoist(PG_PROBLEM_CONSTS, %d, T(%r));" % (i, i, float(h)) for i, h in enumerate(hoisted)]
        L += ["        }", "    };"]
        L += _fn("void f(const T (&x)[N], const T (&u)[M], T (&dx)[N])",
                 emit_body([("dx[%d]" % i, s.f[i]) for i in range(n)], names, hoisted=hoisted))
        # anchored trig (see anchored_trig_args): argument values, their increment between two states, and f
        # with sin/cos given
        targs = anchored_trig_args(list(s.f), xs)
        nt = len(targs)
        L.insert(L.index("struct Problem {") + 1, "    static constexpr int NTRIG = %d, NTRIG_ = %d;" % (nt, max(1, nt)))
        xren = {x: sp.Symbol("x[%d]" % i) for i, x in enumerate(xs)}
        kren = {x: sp.Symbol("k[%d]" % i) for i, x in enumerate(xs)}
        L += _fn("void trig_arg(const T (&x)[N], T (&a)[NTRIG_])",
                 ["(void)x; (void)a;"] + ["a[%d] = %s;" % (j, _cxx(a.xreplace(xren))) for j, (a, _, _) in enumerate(targs)])
        # increment of every argument when the state moves by w*k (the arguments are affine: exact)
        L += _fn("void trig_inc(const T (&k)[N], const T w, T (&d)[NTRIG_])",
                 ["(void)k; (void)w; (void)d;"] +
                 ["d[%d] = w*(%s);" % (j, _cxx(sum(sp.diff(a, x) * kren[x] for x in xs if sp.diff(a, x) != 0)))
                  for j, (a, _, _) in enumerate(targs)])
        if nt:
            body = emit_body([("dx[%d]" % i, s.f[i]) for i in range(n)], names, hoisted=hoisted, given_trig=True)
        else:
            body = ["(void)sn; (void)cs;", "f(m, x, u, dx);"]
        L += _fn("void f_trig(const T (&x)[N], const T (&u)[M], const T (&sn)[NTRIG_], const T (&cs)[NTRIG_], T (&dx)[N])", body)
        ab_out = [("A[%d]" % (i * n + j), s.A[i, j]) for i in range(n) for j in range(n)]
        ab_out += [("B[%d]" % (i * m + j), s.B[i, j]) for i in range(n) for j in range(m)]
        L += _fn("void AB(const T (&x)[N], const T (&u)[M], T (&A)[N * N], T (&B)[N * M])", emit_body(ab_out, names))
        body = emit_body([("const T c", self.cost)], names)
        L += _fn("T cost(const T (&x)[N], const T (&u)[M], const T (&xn)[N], const T t)",
                 ["(void)x; (void)u; (void)xn; (void)t;"] + body + ["return c;"])
        body = emit_body([("const T c", self.fcost)], names)
        L += _fn("T fcost(const T (&x)[N])", ["(void)x;"] + body + ["return c;"])
        if cx is not None:
            outs = [("cx[%d]" % i, cx[0, i]) for i in range(n)] + [("cu[%d]" % j, cu[0, j]) for j in range(m)]
            outs += [("Cxx[%d]" % (i * n + j), Cxx[i, j]) for i in range(n) for j in range(n)]
            outs += [("Cuu[%d]" % (i * m + j), Cuu[i, j]) for i in range(m) for j in range(m)]
            outs += [("Cxu[%d]" % (i * m + j), Cxu[i, j]) for i in range(n) for j in range(m)]
            body = ["(void)x; (void)u; (void)t;"] + emit_body(outs, names)
        else:
            body = ["(void)x; (void)u; (void)t; (void)cx; (void)cu; (void)Cxx; (void)Cuu; (void)Cxu;"]
        L += _fn("void cost_derivs(const T (&x)[N], const T (&u)[M], const T t, T (&cx)[N], T (&cu)[M], "
                 "T (&Cxx)[N * N], T (&Cuu)[M * M], T (&Cxu)[N * M])", body)
        outs = [("cfx[%d]" % i, cfx[0, i]) for i in range(n)]
        outs += [("Cfxx[%d]" % (i * n + j), Cfxx[i, j]) for i in range(n) for j in range(n)]
        L += _fn("void fcost_derivs(const T (&x)[N], T (&cfx)[N], T (&Cfxx)[N * N])",
                 ["(void)x;"] + emit_body(outs, names))
        conds = " || ".join("pg::abs(x[%d]) > T(%r)" % (i, l) for i, l in self.term) or "false"
        L += _fn("bool terminate(const T (&x)[N])", ["(void)x;", "return %s;" % conds])
        L.append("};")
        return "\n".join(L) + "\n"
