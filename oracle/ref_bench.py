"""The UNMODIFIED reference timed on the host: BASELINE.json configs[0] -- examples/ilqr cart-pole swing-up,
single trajectory, horizon 100, the reference's own CPU path (scipy solve_ivp RK45 over lambdified sympy dynamics).
Known answer: 62 iterations, final cost 72.58216861737493.  BASELINE TOOLING ONLY (bench.py's cpu_baseline /
--impl reference legs); needs oracle/_ref (oracle/install_ref.py) or /root/reference."""
import time

import numpy as np

from . import ref_harness as rh

C1_ITERATIONS, C1_COST = 62, 72.58216861737493


def available():
    return rh.reference_root() is not None


def run_c1(max_iters=500):
    """Returns dict(iterations, cost, seconds, env_steps): one `iLQR.run_optim()` of the reference on C1."""
    pg = rh.load_reference()

    def c_k(x, u):  # examples/nmpc/cart_pole.py:7-11
        x1, x2, x3, x4 = x
        u1, = u
        return .5 * (20. * x1 ** 2 + 20 * x2 ** 2 + .01 * x3 ** 2 + .01 * x4 ** 2) + 0.01 * u1 ** 2

    def c_N(x):  # examples/ilqr/cart_pole.py:14-17
        x1, x2, x3, x4 = x
        return 100 * x1 ** 2 + 100 * x2 ** 2 + 10 * x3 ** 2 + 10 * x4 ** 2
    with rh.quiet():
        env = pg.environments.CartPole(c_k, [0, np.pi, 0, 0], 0.02)
        alg = pg.algorithms.ilqr.iLQR(env, 2.0, 0.02, fcost=c_N, constrained=False, path=None, printing=False,
                                      maxIters=max_iters, final_backpass=False)
        iters = [0]
        orig = alg.backward_pass

        def counted(*a, **k):
            iters[0] += 1
            return orig(*a, **k)
        alg.backward_pass = counted
        t0 = time.perf_counter()
        xx, uu, cost = alg.run_optim()
        secs = time.perf_counter() - t0
    return {"iterations": iters[0], "cost": float(cost), "seconds": secs}
