"""Learned dynamics model (reference: pygent/nn_models.py:152-229 `NNDynamics`).

Same constructor, attributes (`layer1/2/3`, `xMean ... dxVar`) and methods (`forward`, `ode`, `save_moments`,
`load_moments`); parameters live in a torch module so that the reference's training loop (Adagrad + MSE,
pygent/algorithms/mbrl.py:450-469) and checkpoints (`state_dict`) work unchanged.  `ode` accepts a batch of
states and runs on the GPU through `pygent_b200.mlp` (there is no CPU evaluation path in this package).
"""
import pickle

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HIDDEN = 500


class NNDynamics(nn.Module):
    def __init__(self, xDim, uDim, dxDim=None, oDim=None, xIsAngle=None):
        super(NNDynamics, self).__init__()
        self.xDim = xDim
        self.oDim = xDim if oDim is None else oDim
        self.dxDim = xDim if dxDim is None else int(dxDim)
        self.uDim = uDim
        self.xIsAngle = [False] * xDim if xIsAngle is None else list(xIsAngle)
        self.uMean = torch.zeros((1, self.uDim))
        self.uVar = torch.ones((1, self.uDim))
        self.oMean = torch.zeros((1, self.oDim))
        self.oVar = torch.ones((1, self.oDim))
        self.xMean = torch.zeros((1, self.xDim))
        self.xVar = torch.ones((1, self.xDim))
        self.dxMean = torch.zeros((1, self.dxDim))
        self.dxVar = torch.ones((1, self.dxDim))
        self.layer1 = nn.Linear(self.oDim + self.uDim, HIDDEN)
        self.layer2 = nn.Linear(HIDDEN, HIDDEN)
        self.layer3 = nn.Linear(HIDDEN, self.dxDim)
        self._device_model = None
        self._weights_version = 0

    def forward(self, o, u):
        h1 = F.relu(self.layer1(torch.cat((o, u), 1)))
        h2 = F.relu(self.layer2(h1))
        return self.layer3(h2)

    def invalidate(self):
        """Call after changing weights or moments: the packed device copy is rebuilt on next use (and the training kernels
        rebuild their operand images of W2: `DynamicsTrainer` watches the version counter)."""
        self._device_model = None
        self._weights_version = getattr(self, "_weights_version", 0) + 1

    def device_model(self, device=None):
        from . import mlp
        if self._device_model is None:
            self._device_model = mlp.MLPDynamics.from_module(self, device=device)
        return self._device_model

    def ode(self, x, u, dt=1.0, precision="fp32"):
        """Network output dx = nn(obs(x), u) * dxVar + dxMean for one state (`[n]`, returns `[dxDim]`, like the
        reference's nn_models.py:193-202) or a batch (`[B, n]` -> `[B, dxDim]`)."""
        x, u = np.asarray(x, dtype=float), np.asarray(u, dtype=float)
        single = x.ndim == 1
        rhs = self.device_model().ode(np.atleast_2d(x), np.atleast_2d(u), dt=1.0, precision=precision)
        dx = rhs[:, self.xDim - self.dxDim:]  # with dt = 1 the velocity half of MBRL.ode is the network output
        return dx[0] if single else dx

    def save_moments(self, filename, path):
        d = {k: getattr(self, k) for k in ("xMean", "xVar", "uMean", "uVar", "oMean", "oVar", "dxMean", "dxVar")}
        for p in (path + 'data/moments_dict.p', path + 'data/moments_dict' + filename + '.p'):
            with open(p, 'wb') as fh:
                pickle.dump(d, fh)

    def load_moments(self, path):
        with open(path + 'data/moments_dict.p', 'rb') as fh:
            d = pickle.load(fh)
        for k, v in d.items():
            setattr(self, k, v)
        self.invalidate()
