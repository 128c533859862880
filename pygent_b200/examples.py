"""The reference's example problems (`examples/ilqr/*.py`, `examples/nmpc/cart_pole.py`) as ready-made cases.

Each cost factory returns (c_k, c_N, params): the python callables a user hands to the environments
and to iLQR, and the same cost written in the quadratic family
    c = sum qx_i x_i^2 + sum (ru_j + re_j exp(-rd t)) u_j^2,   cf = sum qf_i x_i^2
(used by the test oracle and by bench.py's CPU baseline).  `make_env` / `make_ilqr` build the
environment and solver of a named case; `__graft_entry__.build()` precompiles their CUDA libraries.
"""
import numpy as np


def cart_pole_nmpc():  # examples/nmpc/cart_pole.py:7-11 ; final cost examples/ilqr/cart_pole.py:14-17
    def c_k(x, u):
        x1, x2, x3, x4 = x
        u1, = u
        return .5 * (20. * x1 ** 2 + 20 * x2 ** 2 + .01 * x3 ** 2 + .01 * x4 ** 2) + 0.01 * u1 ** 2

    def c_N(x):
        x1, x2, x3, x4 = x
        return 100 * x1 ** 2 + 100 * x2 ** 2 + 10 * x3 ** 2 + 10 * x4 ** 2
    return c_k, c_N, dict(qx=[10, 10, .005, .005], ru=[0.01], qf=[100, 100, 10, 10])


def cart_pole_tv():  # examples/ilqr/cart_pole.py:7-17
    def c_k(x, u, t, mod):
        x1, x2, x3, x4 = x
        u1, = u
        return 15 * x1 ** 2 + 5 * x2 ** 2 + 0.01 * x3 ** 2 + 0.01 * x4 ** 2 + 0.05 * (1000 * mod.exp(-t * 10) + 1) * u1 ** 2

    def c_N(x):
        x1, x2, x3, x4 = x
        return 100 * x1 ** 2 + 100 * x2 ** 2 + 10 * x3 ** 2 + 10 * x4 ** 2
    return c_k, c_N, dict(qx=[15, 5, .01, .01], ru=[0.05], re=[50.0], rd=10.0, qf=[100, 100, 10, 10])


def acrobot():  # examples/ilqr/acrobot.py:7-17
    def c_k(x, u):
        x1, x2, x3, x4 = x
        u1, = u
        return 5 * x1 ** 2 + x2 ** 2 + 0.01 * x3 ** 2 + 0.01 * x4 ** 2 + 0.005 * u1 ** 2

    def c_N(x):
        x1, x2, x3, x4 = x
        return 100 * x1 ** 2 + 100 * x2 ** 2 + 10 * x3 ** 2 + 10 * x4 ** 2
    return c_k, c_N, dict(qx=[5, 1, .01, .01], ru=[0.005], qf=[100, 100, 10, 10])


def double_pendulum():  # examples/ilqr/cart_pole_double_parallel.py:7-17
    def c_k(x, u):
        x1, x2, x3, x4, x5, x6 = x
        u1, = u
        return 0.5 * (2 * x1 ** 2 + 10. * x2 ** 2 + 10. * x3 ** 2 + 0.5 * x4 ** 2 + .1 * u1 ** 2)

    def c_N(x):
        x1, x2, x3, x4, x5, x6 = x
        return 0.5 * (100. * x1 ** 2 + 100. * x2 ** 2 + 100. * x3 ** 2)
    return c_k, c_N, dict(qx=[1, 5, 5, .25, 0, 0], ru=[0.05], qf=[50, 50, 50, 0, 0, 0])


def pendulum():  # examples/ilqr/pendulum.py:7-16
    def c_k(x, u, t, mod):
        x1, x2 = x
        u1, = u
        return x1 ** 2 + .01 * x2 ** 2 + 0.5 * (1000 * mod.exp(-t * 15) + 1) * u1 ** 2

    def c_N(x):
        x1, x2 = x
        return 100 * x1 ** 2 + 1 * x2 ** 2
    return c_k, c_N, dict(qx=[1, .01], ru=[0.5], re=[500.0], rd=15.0, qf=[100, 1])


def triple():
    def c_k(x, u):
        return sum((i + 1) * 0.5 * xi ** 2 for i, xi in enumerate(x)) + 0.01 * u[0] ** 2

    def c_N(x):
        return sum(10 * xi ** 2 for xi in x)
    return c_k, c_N, dict(qx=[0.5 * (i + 1) for i in range(8)], ru=[0.01], qf=[10] * 8)


def car():  # examples/ilqr/car.py:7-17 (m = 2)
    def c_k(x, u):
        x1, x2, x3, x4 = x
        u1, u2 = u
        return x1 ** 2 + x2 ** 2 + 10 * u1 ** 2 + 0.1 * u2 ** 2

    def c_N(x):
        x1, x2, x3, x4 = x
        return 100 * x1 ** 2 + 100 * x2 ** 2 + 1000 * x3 ** 2 + 300 * x4 ** 2
    return c_k, c_N, dict(qx=[1, 1, 0, 0], ru=[10, 0.1], qf=[100, 100, 1000, 300])


def marbot():  # weights of examples/ddpg/marbot.py:6-10, in the iLQR convention c(x, u)
    def c_k(x, u):
        x1, x2, x3, x4 = x
        u1, = u
        return 8.7 * x1 ** 2 + 8.7 * x2 ** 2 + 10e-5 * x3 ** 2 + 10e-5 * x4 ** 2 + 4.7 * u1 ** 2

    def c_N(x):
        x1, x2, x3, x4 = x
        return 100 * x1 ** 2 + 100 * x2 ** 2 + 10 * x3 ** 2 + 10 * x4 ** 2
    return c_k, c_N, dict(qx=[8.7, 8.7, 10e-5, 10e-5], ru=[4.7], qf=[100, 100, 10, 10])


def angular_pendulum():  # examples/ilqr/angular_pendulum.py:7-16 (affine in x1)
    def c_k(x, u):
        x1, x2, x3 = x
        u1, = u
        return (1 - x1) + 0.1 * x3 ** 2 + 0.01 * u1 ** 2

    def c_N(x):
        x1, x2, x3 = x
        return 100 * (1 - x1) + 0.1 * x3 ** 2
    return c_k, c_N, dict(qx=[0, 0, 0.1], ru=[0.01], qf=[0, 0, 0.1], lx=[-1, 0, 0], c0=1.0, lf=[-100, 0, 0], cf0=100.0)


# golden case -> (env class name, env kwargs, system name, cost factory, nominal x0, dt, terminate box)
CASES = {
    "cart_pole": ("CartPole", {}, "cart_pole_lin", cart_pole_nmpc, [0, np.pi, 0, 0], 0.02, {0: 1.2, 2: 4.0}),
    "cart_pole_tv": ("CartPole", {}, "cart_pole_lin", cart_pole_tv, [0, np.pi, 0, 0], 0.02, {0: 1.2, 2: 4.0}),
    "cart_pole_force": ("CartPole", {"linearized": False}, "cart_pole", cart_pole_nmpc, [0, np.pi, 0, 0], 0.02, {0: 1.2, 2: 4.0}),
    "acrobot": ("Acrobot", {}, "acrobot_lin", acrobot, [np.pi, 0., 0., 0.], 0.03, {2: 50., 3: 50.}),
    "acrobot_torque": ("Acrobot", {"linearized": False}, "acrobot", acrobot, [np.pi, 0., 0., 0.], 0.03, {2: 50., 3: 50.}),
    "double_parallel": ("CartPoleDoubleParallel", {}, "cart_pole_double_parallel_lin", double_pendulum,
                        [0, np.pi, np.pi, 0, 0, 0], 0.01, {0: 1.5, 4: 25., 5: 25.}),
    "double_parallel_force": ("CartPoleDoubleParallel", {"linearized": False}, "cart_pole_double_parallel",
                              double_pendulum, [0, np.pi, np.pi, 0, 0, 0], 0.01, {0: 1.5, 4: 25., 5: 25.}),
    "double_serial": ("CartPoleDoubleSerial", {}, "cart_pole_double_serial_lin", double_pendulum,
                      [0, np.pi, np.pi, 0, 0, 0], 0.01, {0: 1.5, 4: 25., 5: 25.}),
    "double_serial_force": ("CartPoleDoubleSerial", {"linearized": False}, "cart_pole_double_serial", double_pendulum,
                            [0, np.pi, np.pi, 0, 0, 0], 0.01, {0: 1.5, 4: 25., 5: 25.}),
    "triple": ("CartPoleTriple", {}, "cart_pole_triple_lin", triple, [0, np.pi, np.pi, np.pi, 0, 0, 0, 0], 0.005, {0: 1.5}),
    "pendulum": ("Pendulum", {}, "pendulum", pendulum, [np.pi, 0], 0.05, {1: 10., 0: 8 * np.pi}),
    "car": ("Car", {}, "car", car, [1., 1., 1.5 * np.pi, 0.], 0.03, {0: 5.0, 1: 5.0}),
    "marbot": ("MarBot", {}, "marbot", marbot, [0., 0.3, 0., 0.], 0.02, {0: 1.0}),
    "angular_pendulum": ("AngularPendulum", {}, "angular_pendulum", angular_pendulum, [-np.cos(0.3), np.sin(0.3), 0.], 0.05, {2: 10.0}),
}

# golden iLQR run -> (case, T, S, integrator (0 RK4, 1 Euler), iLQR kwargs)
ILQR_RUNS = {
    "ilqr_cart_pole_rk4x4": ("cart_pole", 100, 4, 0, {}),
    "ilqr_cart_pole_rk4x1": ("cart_pole", 100, 1, 0, {}),
    "ilqr_cart_pole_parallel_rk4x4": ("cart_pole", 100, 4, 0, {"parallel": True}),
    "ilqr_cart_pole_regtype1_rk4x4": ("cart_pole", 100, 4, 0, {"regType": 1}),
    "ilqr_cart_pole_tv_rk4x4": ("cart_pole_tv", 100, 4, 0, {}),
    "ilqr_acrobot_parallel_rk4x4": ("acrobot", 100, 4, 0, {"parallel": True}),
    "ilqr_double_parallel_rk4x4": ("double_parallel", 100, 4, 0, {}),
    "ilqr_double_parallel_parallel_rk4x4": ("double_parallel", 100, 4, 0, {"parallel": True}),
    "ilqr_double_serial_rk4x4": ("double_serial", 100, 4, 0, {}),
    "ilqr_pendulum_fd_rk4x4": ("pendulum", 60, 4, 0, {}),
    "ilqr_cart_pole_euler": ("cart_pole", 100, 4, 1, {"fastForward": True}),
    "ilqr_car_rk4x4": ("car", 100, 4, 0, {}),   # oracle/make_golden_zoo.py
    "ilqr_angular_pendulum_rk4x4": ("angular_pendulum", 100, 4, 0, {}),
    # finite_diff=True: the cost expansion by central differences (helpers.py:41-128)
    "ilqr_cart_pole_costfd_rk4x4": ("cart_pole", 100, 4, 0, {"finite_diff": True}),
    "ilqr_cart_pole_tv_costfd_rk4x4": ("cart_pole_tv", 100, 4, 0, {"finite_diff": True}),
    "ilqr_pendulum_fd_costfd_rk4x4": ("pendulum", 60, 4, 0, {"finite_diff": True}),
    # constrained=True: the reference's box-QP branch (ilqr.py:303-327) with the QP solved by the generator's BVLS stand-in
    # for cvxopt (oracle/ref_harness.install_qp_standin); m = 1 exact clip, m = 2 (car) active-set enumeration here
    "ilqr_cart_pole_constrained_rk4x4": ("cart_pole", 100, 4, 0, {"constrained": True}),
    "ilqr_cart_pole_tv_constrained_rk4x4": ("cart_pole_tv", 100, 4, 0, {"constrained": True}),
    "ilqr_double_parallel_constrained_rk4x4": ("double_parallel", 100, 4, 0, {"constrained": True}),
    "ilqr_pendulum_fd_constrained_rk4x4": ("pendulum", 60, 4, 0, {"constrained": True}),
    "ilqr_car_constrained_rk4x4": ("car", 100, 4, 0, {"constrained": True}),
}




def double_nmpc():  # examples/nmpc/double.py:7-11: regulation around the hanging configuration x2 = x3 = pi
    def c_k(x, u):
        x1, x2, x3, x4, x5, x6 = x
        u1, = u
        return .5 * (20. * x1 ** 2 + 20 * (x2 - np.pi) ** 2 + .01 * x4 ** 2 + .01 * x5 ** 2) + 0.01 * u1 ** 2

    def c_N(x):  # the example passes none (fcost=None does not run under sympy 1.14); oracle/make_golden_nmpc.py uses this one
        x1, x2, x3, x4, x5, x6 = x
        return 100 * x1 ** 2 + 100 * (x2 - np.pi) ** 2 + 100 * (x3 - np.pi) ** 2 + 10 * x4 ** 2 + 10 * x5 ** 2 + 10 * x6 ** 2
    return c_k, c_N


def triple_nmpc():  # examples/nmpc/triple.py:7-11
    def c_k(x, u):
        x1, x2, x3, x4, x5, x6, x7, x8 = x
        u1, = u
        return .5 * (20. * x1 ** 2 + 20 * (x2 - np.pi) ** 2 + .01 * x5 ** 2 + .01 * x6 ** 2) + 0.01 * u1 ** 2

    def c_N(x):
        x1, x2, x3, x4, x5, x6, x7, x8 = x
        return (100 * x1 ** 2 + 100 * (x2 - np.pi) ** 2 + 100 * (x3 - np.pi) ** 2 + 100 * (x4 - np.pi) ** 2 + 10 * x5 ** 2
                + 10 * x6 ** 2 + 10 * x7 ** 2 + 10 * x8 ** 2)
    return c_k, c_N


# the reference's receding-horizon examples beyond the cart-pole: name -> (env class, cost factory, x0, dt)
NMPC_EXAMPLES = {
    "double": ("CartPoleDoubleSerial", double_nmpc, [0.5, np.pi, np.pi, 0, 0, 0], 0.01),
    "triple": ("CartPoleTriple", triple_nmpc, [0.5, np.pi, np.pi, np.pi, 0, 0, 0, 0], 0.01),
}


def make_nmpc_example(name, x0=None, horizon=1.0, **kw):
    """NMPC of examples/nmpc/double.py / triple.py: plants x0 [B, n] or [n] (default: the example's start)."""
    from . import environments
    from .algorithms.nmpc import NMPC
    cls, costf, x0_default, dt = NMPC_EXAMPLES[name]
    c_k, c_N = costf()
    x0 = x0_default if x0 is None else x0
    env, mpc_env = getattr(environments, cls)(c_k, x0, dt), getattr(environments, cls)(c_k, x0, dt)
    kw.setdefault("ilqr_print", False)
    return NMPC(env, mpc_env, 1.0, dt, horizon, fcost=c_N, constrained=False, step_iterations=1, **kw)


def make_env(case, x0=None, **kw):
    """Environment of a named case; x0 defaults to the case's nominal start (single plant)."""
    from . import environments
    cls, ekw, _, costf, x0_default, dt, _ = CASES[case]
    c_k, _, _ = costf()
    ekw = dict(ekw)
    ekw.update(kw)
    return getattr(environments, cls)(c_k, x0_default if x0 is None else x0, dt, **ekw)


def make_ilqr(case, T, x0=None, env_kw=None, **kw):
    """iLQR solver of a named case with the case's final cost; kw are iLQR kwargs."""
    from .algorithms.ilqr import iLQR
    env = make_env(case, x0, **(env_kw or {}))
    _, c_N, _ = CASES[case][3]()
    dt = CASES[case][5]
    kw.setdefault("printing", False)
    kw.setdefault("final_backpass", False)  # the goldens of the plain solves were made without it (oracle/make_golden.py)
    return iLQR(env, T * dt, dt, fcost=c_N, **kw)


def user_ode_env(x0, dt=0.05):
    """A user-defined `StateSpaceModel` whose ODE is a plain numpy callable (as `Pendulum.ode` is in the
    reference, environments.py:338-349): traced into sympy, compiled, run on the GPU."""
    from .environments import StateSpaceModel

    def ode(t, x, u):
        x1, x2 = x
        u1, = u
        return np.array([x2, u1 + 9.81 * np.sin(x1) - 0.02 * x2])
    return StateSpaceModel(ode, pendulum()[0], x0, 1, dt, name="user_pendulum")
