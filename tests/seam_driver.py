"""Runs in a subprocess (tests/test_gpu_seam.py, __graft_entry__.build): imports the UNMODIFIED reference (oracle/_ref or
/root/reference, through oracle/ref_harness.py's stubs for gym/matplotlib/cvxopt and its pickle shim) with
pygent_b200/shims on sys.path, so that the reference's own `from sympy_to_c import sympy_to_c as sp2c` binds the B200
implementation, then exercises the two call sites of the seam: `CartPole` (modeling_scripts/cart_pole.py:159-170) and
`iLQR.cost_init` (algorithms/ilqr.py:558-571).  Prints one JSON line.   python tests/seam_driver.py <tmpdir> [--compile-only]"""
import contextlib
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pygent_b200", "shims")]

import numpy as np  # noqa: E402

from oracle import ref_harness as rh  # noqa: E402


def main():
    tmp = sys.argv[1]
    compile_only = "--compile-only" in sys.argv
    import sympy_to_c
    from sympy_to_c import sympy_to_c as sp2c
    assert sp2c.convert_to_c.__module__ == "pygent_b200.sympy_to_c"
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        pg = rh.load_reference()

        def c_k(x, u):  # examples/nmpc/cart_pole.py:7-11
            x1, x2, x3, x4 = x
            u1, = u
            return .5 * (20. * x1 ** 2 + 20 * x2 ** 2 + .01 * x3 ** 2 + .01 * x4 ** 2) + 0.01 * u1 ** 2

        def c_N(x):
            x1, x2, x3, x4 = x
            return 100 * x1 ** 2 + 100 * x2 ** 2 + 10 * x3 ** 2 + 10 * x4 ** 2
        env = pg.environments.CartPole(c_k, [0, np.pi, 0, 0], 0.02)
        cwd = os.getcwd()
        os.chdir(tmp)  # iLQR copies sys.argv[0] and creates its result directories relative to `path`
        try:
            alg = pg.algorithms.ilqr.iLQR(env, 0.2, 0.02, fcost=c_N, constrained=False, path=tmp + "/", printing=False, maxIters=1,
                                          final_backpass=False, saving=False)
        except Exception as e:  # noqa: BLE001
            alg, err = None, "%s: %s" % (type(e).__name__, e)
        else:
            err = None
        finally:
            os.chdir(cwd)
    log = buf.getvalue()
    out = {"log_has_c_function": "Using C-function" in log, "log_has_lambdify": "Using lambdify" in log,
           "log_cost_c": "Using C functions for taylor expansion" in log, "ilqr_error": err}
    if not compile_only:
        x, u = [0.1, 0.7, -0.3, 1.1], [2.5]
        out["f"] = np.asarray(env.ode(0, x, u), dtype=float).tolist()
        out["A"] = np.asarray(env.A(x, u), dtype=float).tolist()
        out["B"] = np.asarray(env.B(x, u), dtype=float).ravel().tolist()
        if alg is not None:
            out["cx"] = np.asarray(alg.cx(x, u, 0.3), dtype=float).ravel().tolist()
            out["cu"] = np.asarray(alg.cu(x, u, 0.3), dtype=float).ravel().tolist()
            out["Cxx"] = np.asarray(alg.Cxx(x, u, 0.3), dtype=float).tolist()
            out["Cuu"] = np.asarray(alg.Cuu(x, u, 0.3), dtype=float).tolist()
            out["cfx"] = np.asarray(alg.cfx(x), dtype=float).ravel().tolist()
            out["Cfxx"] = np.asarray(alg.Cfxx(x), dtype=float).tolist()
            out["cost0"] = float(alg.cost)
        # batched evaluation through the same callable: 1,000 points in one launch
        from pygent_b200 import sympy_to_c as s2c
        import sympy as sp
        xs = sp.symbols("x1:5")
        uu = sp.Symbol("u")
        expr = sp.Matrix([xs[2], xs[3], uu, 1.477886191760791 * uu * sp.cos(xs[1]) - 0.06319450526270756 * xs[3] + 14.49806354117336 * sp.sin(xs[1])])
        fn = s2c.convert_to_c((*xs, uu), expr)
        rng = np.random.default_rng(0)
        P = rng.uniform(-2, 2, (5, 1000))
        got = fn(*P)
        want = np.stack([P[2], P[3], P[4], 1.477886191760791 * P[4] * np.cos(P[1]) - 0.06319450526270756 * P[3] + 14.49806354117336 * np.sin(P[1])])
        out["batched_shape"] = list(got.shape)
        out["batched_err"] = float(np.max(np.abs(got[:, 0, :] - want)))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
