"""Goldens of the two hand-written systems of the reference that the example set does not cover:
Car (m = 2 inputs; environments.py:730-761, cost of examples/ilqr/car.py:7-17), MarBot (environments.py:826-868) and
AngularPendulum (environments.py:394-421, the affine cost of examples/ilqr/angular_pendulum.py:7-16).

    python -m oracle.make_golden_zoo        (only works where /root/reference exists)

Same recipe as oracle/make_golden.py (whose helpers are used): control intervals with random actions through the
reference's StateSpaceModel.step (native RK45, RK4 x 1, RK4 x 4, Euler) -> tests/golden/steps_zoo.npz, and one
instrumented iLQR solve of the Car and of the AngularPendulum -> tests/golden/ilqr_<case>_rk4x4.npz.  These classes have
no analytic Jacobians, so the reference linearises them with helpers.system_linearization (central differences, eps = 1e-3).  TEST TOOLING ONLY.
"""
import os
import sys
import time

import numpy as np

from . import make_golden as mg


def _car():  # examples/ilqr/car.py:7-17
    def c_k(x, u):
        x1, x2, x3, x4 = x
        u1, u2 = u
        return x1 ** 2 + x2 ** 2 + 10 * u1 ** 2 + 0.1 * u2 ** 2

    def c_N(x):
        x1, x2, x3, x4 = x
        return 100 * x1 ** 2 + 100 * x2 ** 2 + 1000 * x3 ** 2 + 300 * x4 ** 2
    return c_k, c_N, dict(qx=[1, 1, 0, 0], ru=[10, 0.1], qf=[100, 100, 1000, 300])


def _marbot():  # weights of examples/ddpg/marbot.py:6-10 (there on the next state; here the iLQR convention c(x, u))
    def c_k(x, u):
        x1, x2, x3, x4 = x
        u1, = u
        return 8.7 * x1 ** 2 + 8.7 * x2 ** 2 + 10e-5 * x3 ** 2 + 10e-5 * x4 ** 2 + 4.7 * u1 ** 2

    def c_N(x):
        x1, x2, x3, x4 = x
        return 100 * x1 ** 2 + 100 * x2 ** 2 + 10 * x3 ** 2 + 10 * x4 ** 2
    return c_k, c_N, dict(qx=[8.7, 8.7, 10e-5, 10e-5], ru=[4.7], qf=[100, 100, 10, 10])


def _angular_pendulum():  # examples/ilqr/angular_pendulum.py:7-16 (affine in x1)
    def c_k(x, u):
        x1, x2, x3 = x
        u1, = u
        return (1 - x1) + 0.1 * x3 ** 2 + 0.01 * u1 ** 2

    def c_N(x):
        x1, x2, x3 = x
        return 100 * (1 - x1) + 0.1 * x3 ** 2
    return c_k, c_N, dict(qx=[0, 0, 0.1], ru=[0.01], qf=[0, 0, 0.1], lx=[-1, 0, 0], c0=1.0, lf=[-100, 0, 0], cf0=100.0)


# The example starts the angular pendulum exactly at its hanging equilibrium [-1, 0, 0]: there the first backward pass
# returns k = 0 (the state (cos, sin) decouples from the input to first order), the reference's gradient test fires in
# iteration 1 and nothing is ever optimised.  The golden starts 0.3 rad away from it.
ZOO = {
    "car": ("Car", {}, "car", _car, [1., 1., 1.5 * np.pi, 0.], 0.03),
    "marbot": ("MarBot", {}, "marbot", _marbot, [0., 0.3, 0., 0.], 0.02),
    "angular_pendulum": ("AngularPendulum", {}, "angular_pendulum", _angular_pendulum, [-np.cos(0.3), np.sin(0.3), 0.], 0.05),
}


def main():
    saved = dict(mg.CASES)
    mg.CASES.clear()
    mg.CASES.update(ZOO)
    try:
        rng = np.random.default_rng(20261020)
        if not sys.argv[1:]:
            t0 = time.time()
            np.savez_compressed(os.path.join(mg.GOLDEN, "steps_zoo.npz"), **mg.golden_steps(rng))
            print("steps_zoo.npz %.1fs" % (time.time() - t0))
        # MarBot.ode returns a list: the reference's FD linearisation (helpers.py:144) and fast_step (environments.py:314)
        # fail on it, so the reference itself can only STEP a MarBot (as its DDPG example does) -- no iLQR golden exists
        only = sys.argv[1:]
        from . import ref_harness as rh
        rh.load_reference()
        rh.install_qp_standin()
        for name, case, T, kw in (("ilqr_car_rk4x4", "car", 100, {}), ("ilqr_angular_pendulum_rk4x4", "angular_pendulum", 100, {}),
                                  ("ilqr_car_constrained_rk4x4", "car", 100, {"constrained": True})):  # m = 2 box QP
            if only and name not in only:
                continue
            t0 = time.time()
            g = mg.golden_ilqr(case, T, 4, **kw)
            np.savez_compressed(os.path.join(mg.GOLDEN, name + ".npz"), **g)
            print("%-24s iters=%-4d cost=%.12f  %.1fs" % (name, g["iters"], g["cost"], time.time() - t0))
    finally:
        mg.CASES.clear()
        mg.CASES.update(saved)


if __name__ == "__main__":
    main()
