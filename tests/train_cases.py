"""Shared fixtures of the dynamics-training tests: the deterministic network, the synthetic data set of the golden run
(oracle/make_golden_train.py) and the fixed-batch loop of tests/golden/mlp_training.npz."""
import numpy as np
import torch

from oracle import mlp_oracle as mo
from oracle.make_golden_train import synthetic_transitions
from pygent_b200.data import DataSet
from pygent_b200.nn_models import NNDynamics


def fresh_net():
    w = mo.make_weights(6, 2)
    net = NNDynamics(4, 1, dxDim=2, oDim=5, xIsAngle=[False, True, False, False])
    with torch.no_grad():
        for i, l in enumerate((net.layer1, net.layer2, net.layer3), 1):
            l.weight.copy_(torch.from_numpy(w["W%d" % i]))
            l.bias.copy_(torch.from_numpy(w["b%d" % i]))
    return net


def dataset():
    D = DataSet(10 ** 6)
    for s in synthetic_transitions():
        D.force_add_sample(s)
    return D


def run_fixed_batches(trainer, D, epochs=2):
    trainer.nn_dynamics.eval()
    trainer.set_moments(D)
    trainer.set_data(D)
    losses = []
    for _ in range(epochs):
        for idx in D.batch_indices(512, shuffle=False):
            losses.append(float(trainer.train_step(idx)))
    return losses
