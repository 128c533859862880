"""Dynamics-model training (pygent_b200/algorithms/mbrl.py) against the unmodified reference's training step
(tests/golden/mlp_training.npz, made by oracle/make_golden_train.py), and its data-parallel form on a world_size-2
gloo group.  CPU only: the algebra is library GEMMs through torch; the GPU planner path is tested in test_gpu_mlp.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mlp_oracle as mo
from oracle.make_golden_train import synthetic_transitions
from pygent_b200.algorithms.mbrl import DynamicsTrainer
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
    col = D.columns()
    losses = []
    for _ in range(epochs):
        for idx in D.batch_indices(512, shuffle=False):
            losses.append(float(trainer.train_step({k: v[idx] for k, v in col.items()})))
    return losses


def test_training_step_matches_reference(golden):
    """Moments, standardised MSE and six Adagrad steps (weight decay 1e-3) on fixed batches of 512/512/176."""
    torch.set_num_threads(1)
    g = golden("mlp_training")
    net = fresh_net()
    tr = DynamicsTrainer(net, dyn_lr=1e-3, weight_decay=1e-3)
    losses = run_fixed_batches(tr, dataset())
    for k in ("xMean", "xVar", "dxMean", "dxVar", "oMean", "oVar", "uMean", "uVar"):
        assert np.allclose(getattr(net, k).numpy(), g[k], rtol=1e-6, atol=1e-7), k
    assert np.allclose(losses, g["losses"], rtol=2e-5)
    assert np.allclose(net.layer3.weight.detach().numpy(), g["W3"], rtol=1e-4, atol=1e-6)
    assert np.allclose(net.layer1.weight.detach().numpy(), g["W1"], rtol=1e-4, atol=1e-6)
    assert np.allclose(net.layer1.bias.detach().numpy(), g["b1"], rtol=1e-4, atol=1e-6)
    assert np.allclose(net.layer2.weight.detach().numpy()[:8, :8], g["W2_block"], rtol=1e-4, atol=1e-6)
    assert abs(float(net.layer2.weight.detach().double().sum()) - float(g["W2_sum"])) < 1e-3
    # a single transition's prediction error, as MBRL.pred_loss reports it (mbrl.py:187-206)
    s = dataset().data[3]
    dx_n, f = tr.training_data([s])
    assert abs(tr.pred_loss(s) - float(np.linalg.norm((f - dx_n).detach().numpy(), 2))) < 1e-6


def test_dataset_mirrors_reference_fifo_and_batching():
    D = DataSet(5)
    for i in range(7):
        D.force_add_sample({"i": i})
    assert [s["i"] for s in D.data] == [2, 3, 4, 5, 6]
    assert D.add_sample({"i": 3}) is True and D.add_sample({"i": 9}) is False and len(D.data) == 5
    D2 = dataset()
    sizes = [len(b) for b in D2.batches(512)]
    assert sizes == [512, 512, 176]
    sh = D2.batch_indices(512)
    assert sorted(np.concatenate(sh).tolist()) == list(range(1200))
    assert [len(b) for b in DataSet(10).batches(4)] == []


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        net = fresh_net()
        tr = DynamicsTrainer(net, dyn_lr=1e-3, weight_decay=1e-3)
        losses = run_fixed_batches(tr, dataset(), epochs=1)
        q.put((rank, losses, net.layer3.weight.detach().numpy().copy(), float(net.layer2.weight.detach().double().sum()), tr.collectives))
    finally:
        dist.destroy_process_group()


def test_data_parallel_training_equals_single_process(golden):
    """world_size 2 (gloo): strided batch slices + ONE all-reduce of the flat gradient bucket per step give the
    single-process update (to fp32 summation order), identically on both ranks."""
    g = golden("mlp_training")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.set_num_threads(1)
    net = fresh_net()
    single = run_fixed_batches(DynamicsTrainer(net, dyn_lr=1e-3, weight_decay=1e-3), dataset(), epochs=1)
    for rank, losses, W3, w2sum, ncoll in res:
        assert ncoll == 3  # one collective per optimiser step
        assert np.allclose(losses, single, rtol=1e-5) and np.allclose(losses, g["losses"][:3], rtol=2e-5)
        assert np.allclose(W3, net.layer3.weight.detach().numpy(), rtol=1e-4, atol=1e-6)
        assert abs(w2sum - float(net.layer2.weight.detach().double().sum())) < 1e-3
    assert np.array_equal(res[0][2], res[1][2]) and res[0][3] == res[1][3]  # ranks stay bit-identical
