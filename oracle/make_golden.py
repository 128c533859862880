"""Freeze outputs of the UNMODIFIED reference into tests/golden/*.npz.

    python -m oracle.make_golden          (only works where /root/reference exists)

The fixtures pin the oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.  Every
array below is produced by the reference's own code (pygent.environments / pygent.algorithms.ilqr),
imported through oracle/ref_harness.py; "rk4xS" runs use the reference with only the solve_ivp line
of StateSpaceModel.step replaced by S classical RK4 substeps (the fixed-step parity schedule).
"""
import os
import sys
import time

import numpy as np

from . import ref_harness as rh

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# ---- the example costs of the reference, verbatim semantics ---------------------------------------
# each entry: python callables for the reference + the parameters of the oracle's cost family


def _cart_pole_nmpc():  # examples/nmpc/cart_pole.py:7-11, final cost examples/ilqr/cart_pole.py:14-17
    def c_k(x, u):
        x1, x2, x3, x4 = x
        u1, = u
        return .5 * (20. * x1 ** 2 + 20 * x2 ** 2 + .01 * x3 ** 2 + .01 * x4 ** 2) + 0.01 * u1 ** 2

    def c_N(x):
        x1, x2, x3, x4 = x
        return 100 * x1 ** 2 + 100 * x2 ** 2 + 10 * x3 ** 2 + 10 * x4 ** 2
    return c_k, c_N, dict(qx=[10, 10, .005, .005], ru=[0.01], qf=[100, 100, 10, 10])


def _cart_pole_tv():  # examples/ilqr/cart_pole.py:7-17 (time-weighted control cost)
    def c_k(x, u, t, mod):
        x1, x2, x3, x4 = x
        u1, = u
        return 15 * x1 ** 2 + 5 * x2 ** 2 + 0.01 * x3 ** 2 + 0.01 * x4 ** 2 + 0.05 * (1000 * mod.exp(-t * 10) + 1) * u1 ** 2

    def c_N(x):
        x1, x2, x3, x4 = x
        return 100 * x1 ** 2 + 100 * x2 ** 2 + 10 * x3 ** 2 + 10 * x4 ** 2
    return c_k, c_N, dict(qx=[15, 5, .01, .01], ru=[0.05], re=[50.0], rd=10.0, qf=[100, 100, 10, 10])


def _acrobot():  # examples/ilqr/acrobot.py:7-17
    def c_k(x, u):
        x1, x2, x3, x4 = x
        u1, = u
        return 5 * x1 ** 2 + x2 ** 2 + 0.01 * x3 ** 2 + 0.01 * x4 ** 2 + 0.005 * u1 ** 2

    def c_N(x):
        x1, x2, x3, x4 = x
        return 100 * x1 ** 2 + 100 * x2 ** 2 + 10 * x3 ** 2 + 10 * x4 ** 2
    return c_k, c_N, dict(qx=[5, 1, .01, .01], ru=[0.005], qf=[100, 100, 10, 10])


def _double_parallel():  # examples/ilqr/cart_pole_double_parallel.py:7-17
    def c_k(x, u):
        x1, x2, x3, x4, x5, x6 = x
        u1, = u
        return 0.5 * (2 * x1 ** 2 + 10. * x2 ** 2 + 10. * x3 ** 2 + 0.5 * x4 ** 2 + .1 * u1 ** 2)

    def c_N(x):
        x1, x2, x3, x4, x5, x6 = x
        return 0.5 * (100. * x1 ** 2 + 100. * x2 ** 2 + 100. * x3 ** 2)
    return c_k, c_N, dict(qx=[1, 5, 5, .25, 0, 0], ru=[0.05], qf=[50, 50, 50, 0, 0, 0])


def _double_serial():  # same weights as the parallel example, on the serial pendulum
    return _double_parallel()


def _pendulum():  # examples/ilqr/pendulum.py:7-16
    def c_k(x, u, t, mod):
        x1, x2 = x
        u1, = u
        return x1 ** 2 + .01 * x2 ** 2 + 0.5 * (1000 * mod.exp(-t * 15) + 1) * u1 ** 2

    def c_N(x):
        x1, x2 = x
        return 100 * x1 ** 2 + 1 * x2 ** 2
    return c_k, c_N, dict(qx=[1, .01], ru=[0.5], re=[500.0], rd=15.0, qf=[100, 1])


def _triple():
    def c_k(x, u):
        return sum((i + 1) * 0.5 * xi ** 2 for i, xi in enumerate(x)) + 0.01 * u[0] ** 2

    def c_N(x):
        return sum(10 * xi ** 2 for xi in x)
    return c_k, c_N, dict(qx=[0.5 * (i + 1) for i in range(8)], ru=[0.01], qf=[10] * 8)


# name -> (reference class name, kwargs, oracle/product system name, cost factory, x0, dt, uMax)
CASES = {
    "cart_pole": ("CartPole", {}, "cart_pole_lin", _cart_pole_nmpc, [0, np.pi, 0, 0], 0.02),
    "cart_pole_tv": ("CartPole", {}, "cart_pole_lin", _cart_pole_tv, [0, np.pi, 0, 0], 0.02),
    "cart_pole_force": ("CartPole", {"linearized": False}, "cart_pole", _cart_pole_nmpc, [0, np.pi, 0, 0], 0.02),
    "acrobot": ("Acrobot", {}, "acrobot_lin", _acrobot, [np.pi, 0., 0., 0.], 0.03),
    "acrobot_torque": ("Acrobot", {"linearized": False}, "acrobot", _acrobot, [np.pi, 0., 0., 0.], 0.03),
    "double_parallel": ("CartPoleDoubleParallel", {}, "cart_pole_double_parallel_lin", _double_parallel,
                        [0, np.pi, np.pi, 0, 0, 0], 0.01),
    "double_parallel_force": ("CartPoleDoubleParallel", {"linearized": False}, "cart_pole_double_parallel",
                              _double_parallel, [0, np.pi, np.pi, 0, 0, 0], 0.01),
    "double_serial": ("CartPoleDoubleSerial", {}, "cart_pole_double_serial_lin", _double_serial,
                      [0, np.pi, np.pi, 0, 0, 0], 0.01),
    "double_serial_force": ("CartPoleDoubleSerial", {"linearized": False}, "cart_pole_double_serial", _double_serial,
                            [0, np.pi, np.pi, 0, 0, 0], 0.01),
    "triple": ("CartPoleTriple", {}, "cart_pole_triple_lin", _triple, [0, np.pi, np.pi, np.pi, 0, 0, 0, 0], 0.005),
    "pendulum": ("Pendulum", {}, "pendulum", _pendulum, [np.pi, 0], 0.05),
}


def make_env(case, x0=None):
    pg = rh.load_reference()
    cls, kw, system, costf, x0_default, dt = CASES[case]
    c_k, c_N, cost_params = costf()
    with rh.quiet():
        env = getattr(pg.environments, cls)(c_k, list(x0_default if x0 is None else x0), dt, **kw)
    return env, c_k, c_N, cost_params, system, dt


def golden_models(rng):
    """f, A, B of every system at random points, from the reference's own callables."""
    out = {}
    for case in CASES:
        env, _, _, _, system, _ = make_env(case)
        if system in [k.split("/")[0] for k in out]:
            continue
        n, m = env.xDim, env.uDim
        P = 32
        xs = rng.uniform(-2.5, 2.5, (P, n))
        us = rng.uniform(-3, 3, (P, m))
        out[system + "/x"] = xs
        out[system + "/u"] = us
        out[system + "/f"] = np.array([np.asarray(env.ode(0, list(x), list(u)), dtype=float) for x, u in zip(xs, us)])
        if hasattr(env, "A"):
            out[system + "/A"] = np.array([np.asarray(env.A(list(x), list(u)), dtype=float) for x, u in zip(xs, us)])
            out[system + "/B"] = np.array([np.asarray(env.B(list(x), list(u)), dtype=float).reshape(n, m) for x, u in zip(xs, us)])
    # the two hand-written systems with m=2 / a 2x2 inverse, straight from environments.py
    pg = rh.load_reference()
    for system, cls, n, m in (("car", "Car", 4, 2), ("marbot", "MarBot", 4, 1), ("angular_pendulum", "AngularPendulum", 3, 1)):
        ode = getattr(pg.environments, cls).ode
        xs = rng.uniform(-1.5, 1.5, (32, n))
        us = rng.uniform(-0.5, 0.5, (32, m))
        out[system + "/x"], out[system + "/u"] = xs, us
        out[system + "/f"] = np.array([np.asarray(ode(0, list(x), list(u)), dtype=float) for x, u in zip(xs, us)])
    return out


def golden_steps(rng, T=40):
    """T control intervals with random actions through the reference's StateSpaceModel.step:
    native adaptive RK45, and RK4 x 1 / RK4 x 4; plus fast_step (Euler)."""
    out = {}
    for case in CASES:
        env, _, _, _, system, dt = make_env(case)
        n, m = env.xDim, env.uDim
        x0 = np.array(CASES[case][4], dtype=float) + rng.uniform(-0.05, 0.05, n)
        umax = float(np.asarray(env.uMax).ravel()[0])
        U = rng.uniform(-umax, umax, (T, m)) * 0.5
        out[case + "/x0"], out[case + "/U"] = x0, U
        for tag, S in (("rk45", None), ("rk4x1", 1), ("rk4x4", 4), ("euler", 0)):
            env.reset(list(x0))
            costs = []
            if S is None or S == 0:
                try:
                    for k in range(T):
                        costs.append(env.fast_step(U[k], dt) if S == 0 else env.step(U[k], dt))
                except TypeError:
                    # MarBot.ode returns a list: the reference's own fast_step fails on `dt*list` (environments.py:314);
                    # there is no Euler golden for such a system
                    assert S == 0
                    continue
            else:
                with rh.fixed_step(S):
                    for k in range(T):
                        costs.append(env.step(U[k], dt))
            out["%s/%s/X" % (case, tag)] = np.array(env.history, dtype=float)
            out["%s/%s/c" % (case, tag)] = np.array(costs, dtype=float)
            out["%s/%s/terminated" % (case, tag)] = np.array(env.terminated)
    return out


def golden_ilqr(case, T, S, x0=None, max_iters=500, **kw):
    """One reference iLQR solve, instrumented from the outside (no reference code modified):
    initial trajectory, first backward pass, alpha=1 forward pass, per-iteration trace, result."""
    pg = rh.load_reference()
    iLQR = pg.algorithms.ilqr.iLQR
    env, c_k, c_N, cost_params, system, dt = make_env(case, x0)
    trace = []
    orig_bw = iLQR.backward_pass

    def bw(self, final=False):
        trace.append((float(self.cost), float(self.mu), float(self.current_alpha)))
        return orig_bw(self, final)
    iLQR.backward_pass = bw
    ctx = rh.fixed_step(S) if S else rh.quiet()
    out = {}
    try:
        with ctx, rh.quiet():
            alg = iLQR(env, T * dt, dt, fcost=c_N, constrained=kw.pop("constrained", False), final_backpass=False,
                       printing=False, maxIters=max_iters, **kw)
            assert alg.steps == T, (alg.steps, T)
            out["x0"] = np.array(env.x0, dtype=float)
            out["xx0"], out["uu0"], out["cost0"] = np.array(alg.xx), np.array(alg.uu), float(alg.cost)
            V1, V2, ok, KK, kk = orig_bw(alg)
            out["bw_KK"] = np.array(KK).reshape(T, alg.uDim, alg.xDim)
            out["bw_kk"] = np.array(kk).reshape(T, alg.uDim)
            out["bw_sumV"] = np.array([np.sum(V1), np.sum(V2)])
            xx1, uu1, c1, term = alg.forward_pass(1.0, KK, kk)
            out["fw1_xx"], out["fw1_uu"], out["fw1_cost"] = np.array(xx1), np.array(uu1), float(c1)
            t0 = time.time()
            xx, uu, cost = alg.run_optim()
            out["runtime_s"] = time.time() - t0
    finally:
        iLQR.backward_pass = orig_bw
    out["xx"], out["uu"], out["cost"] = np.array(xx), np.array(uu), float(cost)
    out["KK"] = np.array(alg.KK).reshape(T, alg.uDim, alg.xDim)
    out["kk"] = np.array(alg.kk).reshape(T, alg.uDim)
    out["trace"] = np.array(trace)  # (cost, mu, alpha) at the start of every iteration
    out["iters"] = len(trace)
    out["mu"], out["alpha"] = float(alg.mu), float(alg.current_alpha)
    out["dt"], out["T"], out["S"] = dt, T, S or 0
    return out


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    rng = np.random.default_rng(20261019)
    if not sys.argv[1:]:  # (named runs only: leave the model / step fixtures alone)
        t0 = time.time()
        np.savez_compressed(os.path.join(GOLDEN, "models.npz"), **golden_models(rng))
        print("models.npz  %.1fs" % (time.time() - t0))
        t0 = time.time()
        np.savez_compressed(os.path.join(GOLDEN, "steps.npz"), **golden_steps(rng))
        print("steps.npz   %.1fs" % (time.time() - t0))
    runs = [
        ("ilqr_cart_pole_rk4x4", "cart_pole", 100, 4, {}),
        ("ilqr_cart_pole_rk45", "cart_pole", 100, 0, {}),
        ("ilqr_cart_pole_rk4x1", "cart_pole", 100, 1, {}),
        ("ilqr_cart_pole_parallel_rk4x4", "cart_pole", 100, 4, {"parallel": True}),
        ("ilqr_cart_pole_regtype1_rk4x4", "cart_pole", 100, 4, {"regType": 1}),
        ("ilqr_cart_pole_tv_rk4x4", "cart_pole_tv", 100, 4, {}),
        ("ilqr_acrobot_parallel_rk4x4", "acrobot", 100, 4, {"parallel": True}),
        ("ilqr_double_parallel_rk4x4", "double_parallel", 100, 4, {}),
        ("ilqr_double_parallel_parallel_rk4x4", "double_parallel", 100, 4, {"parallel": True}),
        ("ilqr_double_serial_rk4x4", "double_serial", 100, 4, {}),
        ("ilqr_pendulum_fd_rk4x4", "pendulum", 60, 4, {}),
        ("ilqr_cart_pole_euler", "cart_pole", 100, 0, {"fastForward": True}),
        # finite_diff=True: cost expansion by central differences (helpers.py:41-128, ilqr.py:240-242, :598-603)
        ("ilqr_cart_pole_costfd_rk4x4", "cart_pole", 100, 4, {"finite_diff": True}),
        ("ilqr_cart_pole_tv_costfd_rk4x4", "cart_pole_tv", 100, 4, {"finite_diff": True}),
        ("ilqr_pendulum_fd_costfd_rk4x4", "pendulum", 60, 4, {"finite_diff": True}),
        # constrained=True (ilqr.py:303-327; every shipped examples/ilqr/*.py): the reference's own code around its box QP, the
        # QP itself solved by the BVLS stand-in for the absent cvxopt (ref_harness.install_qp_standin)
        ("ilqr_cart_pole_constrained_rk4x4", "cart_pole", 100, 4, {"constrained": True}),
        ("ilqr_cart_pole_tv_constrained_rk4x4", "cart_pole_tv", 100, 4, {"constrained": True}),
        ("ilqr_double_parallel_constrained_rk4x4", "double_parallel", 100, 4, {"constrained": True}),
        ("ilqr_pendulum_fd_constrained_rk4x4", "pendulum", 60, 4, {"constrained": True}),
    ]
    only = sys.argv[1:]
    rh.load_reference()
    rh.install_qp_standin()
    for name, case, T, S, kw in runs:
        if only and name not in only:
            continue
        t0 = time.time()
        g = golden_ilqr(case, T, S, **dict(kw))
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **g)
        print("%-40s iters=%-4d cost=%.12f  %.1fs" % (name, g["iters"], g["cost"], time.time() - t0))


if __name__ == "__main__":
    main()
