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


if __name__ == "__main__":
    main()
