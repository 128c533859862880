"""Golden closed-loop run of the UNMODIFIED reference's NMPC (pygent/algorithms/nmpc.py: MPCAgent.take_action,
shift_planner) on the cart-pole of examples/nmpc/cart_pole.py, with the fixed-step RK4 x 4 `step` of
oracle/ref_harness.py in place of solve_ivp, constrained=False (cvxopt is absent) and no action noise.
    python -m oracle.make_golden_nmpc"""
import os
import tempfile

import numpy as np

from . import ref_harness as rh
from .make_golden import GOLDEN


def c_k(x, u):
    x1, x2, x3, x4 = x
    u1, = u
    return .5 * (20. * x1 ** 2 + 20 * x2 ** 2 + .01 * x3 ** 2 + .01 * x4 ** 2) + 0.01 * u1 ** 2


def c_N(x):
    x1, x2, x3, x4 = x
    return 100 * x1 ** 2 + 100 * x2 ** 2 + 10 * x3 ** 2 + 10 * x4 ** 2


def c_k_double(x, u):  # examples/nmpc/double.py:7-11 (regulation around the hanging configuration x2 = x3 = pi)
    x1, x2, x3, x4, x5, x6 = x
    u1, = u
    return .5 * (20. * x1 ** 2 + 20 * (x2 - np.pi) ** 2 + .01 * x4 ** 2 + .01 * x5 ** 2) + 0.01 * u1 ** 2


def c_N_double(x):  # the example passes no final cost (fcost=None does not run under sympy 1.14)
    x1, x2, x3, x4, x5, x6 = x
    return 100 * x1 ** 2 + 100 * (x2 - np.pi) ** 2 + 100 * (x3 - np.pi) ** 2 + 10 * x4 ** 2 + 10 * x5 ** 2 + 10 * x6 ** 2


def c_k_triple(x, u):  # examples/nmpc/triple.py:7-11
    x1, x2, x3, x4, x5, x6, x7, x8 = x
    u1, = u
    return .5 * (20. * x1 ** 2 + 20 * (x2 - np.pi) ** 2 + .01 * x5 ** 2 + .01 * x6 ** 2) + 0.01 * u1 ** 2


def c_N_triple(x):
    x1, x2, x3, x4, x5, x6, x7, x8 = x
    return (100 * x1 ** 2 + 100 * (x2 - np.pi) ** 2 + 100 * (x3 - np.pi) ** 2 + 100 * (x4 - np.pi) ** 2 + 10 * x5 ** 2 + 10 * x6 ** 2
            + 10 * x7 ** 2 + 10 * x8 ** 2)


def main():
    import sys
    only = sys.argv[1:]
    if not only or "nmpc_cart_pole" in only:
        run("nmpc_cart_pole", [0, np.pi, 0, 0])                                    # swing-up, inputs saturated throughout
    if not only or "nmpc_cart_pole_stab" in only:
        run("nmpc_cart_pole_stab", [0.2, 0.3, 0.1, -0.2], n_steps=40, horizon=1.0)  # stabilisation around the upright
    if not only or "nmpc_double_serial" in only:   # examples/nmpc/double.py: n = 6, cart displaced by 0.5 m
        run("nmpc_double_serial", [0.5, np.pi, np.pi, 0, 0, 0], n_steps=30, horizon=1.0, dt=0.01, cls="CartPoleDoubleSerial",
            costs=(c_k_double, c_N_double))
    if not only or "nmpc_triple" in only:          # examples/nmpc/triple.py: n = 8
        run("nmpc_triple", [0.5, np.pi, np.pi, np.pi, 0, 0, 0, 0], n_steps=20, horizon=1.0, dt=0.01, cls="CartPoleTriple",
            costs=(c_k_triple, c_N_triple))


def run(name, x0, n_steps=50, horizon=2.0, dt=0.02, init_iterations=500, cls="CartPole", costs=None):
    c_k, c_N = costs or (globals()["c_k"], globals()["c_N"])
    pg = rh.load_reference()
    import pygent.algorithms.ilqr as ref_ilqr
    import pygent.algorithms.nmpc as nmpc
    ref_ilqr.copyfile = lambda *a, **k: None  # the reference copies the executing script next to its results (ilqr.py:110)
    out = {"x0": np.array(x0), "dt": dt, "horizon": horizon, "n_steps": n_steps, "init_iterations": init_iterations}
    with rh.fixed_step(4), rh.quiet(), tempfile.TemporaryDirectory() as tmp:
        env = getattr(pg.environments, cls)(c_k, x0, dt)
        mpc_env = getattr(pg.environments, cls)(c_k, x0, dt)
        agent = nmpc.MPCAgent(mpc_env, horizon, dt, tmp + "/", init_iterations=init_iterations, step_iterations=1, fcost=c_N,
                              constrained=False, fastForward=False, printing=False, saving=False, init_optim=True,
                              finite_diff=False, add_noise=False)
        opt = agent.traj_optimizer
        out.update(init_cost=float(opt.cost), init_xx=np.array(opt.xx, dtype=float), init_uu=np.array(opt.uu, dtype=float),
                   init_env_x=np.array(mpc_env.x, dtype=float),
                   # run_optim ended with the FINAL backward pass (final_backpass=True is the reference's default,
                   # ilqr.py:75, :487-494): these are its gains, and mu is left at 0
                   init_KK=np.array(opt.KK, dtype=float), init_kk=np.array(opt.kk, dtype=float)[:, :, 0], init_mu=float(opt.mu),
                   init_alpha=float(opt.current_alpha))
        env.reset(x0)
        agent.reset()
        costs, plan_costs, accepted, plan_xN = [], [], [], []
        for i in range(n_steps):
            u = agent.take_action(dt, env.x)
            accepted.append(bool(opt.success_fw))
            plan_costs.append(float(opt.cost))
            plan_xN.append(np.array(opt.xx[-1], dtype=float))
            costs.append(float(env.step(u, dt)))
    out.update(X=np.array(env.history, dtype=float), U=np.array(agent.history[1:], dtype=float), costs=np.array(costs),
               plan_costs=np.array(plan_costs), accepted=np.array(accepted), plan_xN=np.array(plan_xN))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name + ".npz: init cost %.6f, closed-loop cost %.6f, accepted %d / %d, x_end %s" % (
        out["init_cost"], float(np.sum(costs)), int(np.sum(accepted)), n_steps, env.history[-1]))


if __name__ == "__main__":
    main()
