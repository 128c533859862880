"""Golden vectors of the learned-dynamics path from the UNMODIFIED reference classes:
pygent.nn_models.NNDynamics (ode), the MBRL.ode formula (mbrl.py:125-130) bound to it, and
pygent.environments.StateSpaceModel.fast_step.   python -m oracle.make_golden_mlp"""
import os

import numpy as np
import torch

from . import mlp_oracle as mo
from . import ref_harness as rh
from .make_golden import GOLDEN


def main():
    pg = rh.load_reference()
    import pygent.nn_models as nnm
    is_angle = [False, True, False, False]  # CartPole (environments.py:465)
    n, m, n_obs, dx = 4, 1, 5, 2
    w, mom = mo.make_weights(n_obs + m, dx), mo.make_moments(n_obs, m, dx)
    net = nnm.NNDynamics(n, m, dxDim=dx, oDim=n_obs, xIsAngle=is_angle)
    with torch.no_grad():
        net.layer1.weight.copy_(torch.from_numpy(w["W1"])); net.layer1.bias.copy_(torch.from_numpy(w["b1"]))
        net.layer2.weight.copy_(torch.from_numpy(w["W2"])); net.layer2.bias.copy_(torch.from_numpy(w["b2"]))
        net.layer3.weight.copy_(torch.from_numpy(w["W3"])); net.layer3.bias.copy_(torch.from_numpy(w["b3"]))
    net.oMean, net.oVar = torch.from_numpy(mom["o_mean"])[None], torch.from_numpy(mom["o_var"])[None]
    net.uMean, net.uVar = torch.from_numpy(mom["u_mean"])[None], torch.from_numpy(mom["u_var"])[None]
    net.dxMean, net.dxVar = torch.from_numpy(mom["dx_mean"])[None], torch.from_numpy(mom["dx_var"])[None]
    dt = 0.02

    def ode(t, x, u):  # mbrl.py:125-128 with sparse_dyn=True
        return np.concatenate((x[int(len(x) / 2):], 1 / dt * net.ode(x, u)))
    rng = np.random.default_rng(7)
    xs = np.array([0, np.pi, 0, 0]) + rng.uniform(-1, 1, (16, 4))
    us = rng.uniform(-8, 8, (16, 1))
    torch.set_num_threads(1)
    rhs = np.array([ode(None, np.array(x), np.array(u)) for x, u in zip(xs, us)])
    with rh.quiet():
        env = pg.environments.StateSpaceModel(ode, lambda x, u: 0.0, list(xs[0]), 1, dt)
    U = rng.uniform(-8, 8, (30, 1))
    env.reset(np.array(xs[0]))
    for k in range(30):
        env.fast_step(U[k], dt)
    np.savez_compressed(os.path.join(GOLDEN, "mlp_cart_pole.npz"), xs=xs, us=us, rhs=rhs, x0=xs[0], U=U,
                        X=np.array(env.history, dtype=float), dt=dt)
    print("mlp_cart_pole.npz", rhs[0], env.history[-1])

    # --- the reference's planner on the learned model: iLQR(StateSpaceModel(MBRL.ode), fastForward=True), which
    # linearises by central differences (ilqr.py:523-527); costs of examples/mbrl/cart_pole_hpc.py:24-33
    def c_k(x, u, mod):
        x1, x2, x3, x4 = x
        u1, = u
        return 10 * x1 ** 2 + 5 * x2 ** 2 + .01 * x3 ** 2 + .01 * x4 ** 2 + 0.1 * u1 ** 2

    def c_N(x, mod):
        x1, x2, x3, x4 = x
        return 100 * x1 ** 2 + 100 * x2 ** 2 + 10 * x3 ** 2 + 10 * x4 ** 2
    x0 = [0.3, np.pi - 0.4, 0.2, -0.5]
    T = 40
    with rh.quiet():
        nn_env = pg.environments.StateSpaceModel(ode, c_k, x0, 1, dt)
        nn_env.uMax = np.array([8.0])
        alg = pg.algorithms.ilqr.iLQR(nn_env, T * dt, dt, fastForward=True, fcost=c_N, constrained=False, path=None,
                                      printing=False, maxIters=12)
    out = {"x0": np.array(x0), "T": T, "dt": dt, "cost0": float(alg.cost), "xx0": np.array(alg.xx, dtype=float)}
    lin = [alg.linearization(x, u) for x, u in zip(alg.xx[:-1], alg.uu)]
    out["Ad0"] = np.array([l[0] for l in lin])
    out["Bd0"] = np.array([l[1] for l in lin])
    V1, V2, ok, KK, kk = alg.backward_pass()
    out.update(KK1=np.array(KK), kk1=np.array(kk).reshape(len(kk), 1), sV1=float(np.sum(V1)), sV2=float(np.sum(V2)), ok1=bool(ok))
    with rh.quiet():
        xx1, uu1, c1, _ = alg.forward_pass(1.0, KK, kk)
    out.update(cost_alpha1=float(c1), xx_alpha1=np.array(xx1, dtype=float))
    with rh.quiet():
        nn_env2 = pg.environments.StateSpaceModel(ode, c_k, x0, 1, dt)
        alg2 = pg.algorithms.ilqr.iLQR(nn_env2, T * dt, dt, fastForward=True, fcost=c_N, constrained=False, path=None,
                                       printing=False, maxIters=12)
        trace = []
        orig = alg2.backward_pass

        def traced(*a, **k):
            trace.append((float(alg2.cost), float(alg2.mu)))
            return orig(*a, **k)
        alg2.backward_pass = traced
        xx, uu, cost = alg2.run_optim()
    out.update(cost=float(cost), xx=np.array(xx, dtype=float), uu=np.array(uu, dtype=float),
               cost_trace=np.array([t[0] for t in trace]), mu_trace=np.array([t[1] for t in trace]))
    np.savez_compressed(os.path.join(GOLDEN, "mlp_ilqr_cart_pole.npz"), **out)
    print("mlp_ilqr_cart_pole.npz cost0 %.6f alpha1 %.6f final %.6f after %d backward passes" % (out["cost0"], out["cost_alpha1"], out["cost"], len(trace)))


if __name__ == "__main__":
    main()
