"""Golden vectors of the dynamics-training step from the UNMODIFIED reference: `MBRL.set_moments` and
`MBRL.training_data` (pygent/algorithms/mbrl.py:471-509) called on a stand-in `self`, `nn.MSELoss`, and
`torch.optim.Adagrad(lr, weight_decay)` exactly as `MBRL.__init__`/`MBRL.training` wire them (:71-73, :450-469), on
fixed batches (no shuffling).     python -m oracle.make_golden_train"""
import os
import types

import numpy as np
import torch

from . import mlp_oracle as mo
from . import ref_harness as rh
from .make_golden import GOLDEN


def synthetic_transitions(n=1200, seed=11):
    """Transitions with the layout MBRL stores (mbrl.py:160-167); values are arbitrary but deterministic."""
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        x_ = np.array([0, np.pi, 0, 0]) + rng.uniform(-1, 1, 4) * np.array([.5, 1.5, 2., 4.])
        u = rng.uniform(-8, 8, 1)
        x = x_ + 0.02 * np.array([x_[2], x_[3], u[0], 1.5 * u[0] * np.cos(x_[1]) + 14.5 * np.sin(x_[1])]) + rng.normal(0, 1e-3, 4)
        obs = lambda s: np.array([s[0], np.cos(s[1]), np.sin(s[1]), s[2], s[3]])
        out.append({"x_": x_, "u": u, "x": x, "o_": obs(x_), "o": obs(x), "c": [0.0], "t": [False]})
    return out


def main():
    pg = rh.load_reference()
    import pygent.nn_models as nnm
    from pygent.algorithms.mbrl import MBRL
    from pygent.data import DataSet
    torch.set_num_threads(1)
    w = mo.make_weights(6, 2)
    net = nnm.NNDynamics(4, 1, dxDim=2, oDim=5, xIsAngle=[False, True, False, False])
    with torch.no_grad():
        for i, l in enumerate((net.layer1, net.layer2, net.layer3), 1):
            l.weight.copy_(torch.from_numpy(w["W%d" % i]))
            l.bias.copy_(torch.from_numpy(w["b%d" % i]))
    stub = types.SimpleNamespace(nn_dynamics=net, sparse_dyn=True)
    D = DataSet(10 ** 6)
    for s in synthetic_transitions():
        D.force_add_sample(s)
    MBRL.set_moments(stub, D)
    optim = torch.optim.Adagrad(net.parameters(), lr=1e-3, weight_decay=1e-3)
    crit = torch.nn.MSELoss()
    net.eval()
    losses = []
    for epoch in range(2):
        for batch in D.batches(512):
            dx, f = MBRL.training_data(stub, batch)
            loss = crit(f, dx)
            optim.zero_grad()
            loss.backward()
            optim.step()
            losses.append(float(loss))
    g = lambda t: t.detach().numpy().copy()
    np.savez_compressed(os.path.join(GOLDEN, "mlp_training.npz"), losses=np.array(losses),
                        xMean=g(net.xMean), xVar=g(net.xVar), dxMean=g(net.dxMean), dxVar=g(net.dxVar), oMean=g(net.oMean),
                        oVar=g(net.oVar), uMean=g(net.uMean), uVar=g(net.uVar), W3=g(net.layer3.weight), b1=g(net.layer1.bias),
                        W2_block=g(net.layer2.weight)[:8, :8], W2_sum=float(net.layer2.weight.double().sum()),
                        W1=g(net.layer1.weight))
    print("mlp_training.npz losses", losses)


if __name__ == "__main__":
    main()
