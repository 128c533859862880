"""Dynamics-model training, CPU side: (1) the torch restatement oracle/train_oracle.py against the unmodified reference's
training step (tests/golden/mlp_training.npz, made by oracle/make_golden_train.py); (2) the product's HOST logic
(pygent_b200/algorithms/mbrl.py `DynamicsTrainer`: moments, batching, strided data-parallel slices, ONE bucket all-reduce
per step) driven through that oracle as an injected backend, single process and on a world_size-2 gloo group.  The product's
own backend is CUDA kernels only (tests/test_gpu_training.py); without a GPU it must refuse to construct."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mlp_oracle as mo
from oracle.make_golden_train import synthetic_transitions
from oracle.train_oracle import TorchBackend
from pygent_b200.algorithms.mbrl import DynamicsTrainer
from pygent_b200.data import DataSet
from pygent_b200.nn_models import NNDynamics  # noqa: F401
from train_cases import dataset, fresh_net, run_fixed_batches


def oracle_trainer(net, **kw):
    return DynamicsTrainer(net, dyn_lr=1e-3, weight_decay=1e-3, backend=TorchBackend(net), **kw)


def test_training_oracle_matches_reference(golden):
    """Moments, standardised MSE and six Adagrad steps (weight decay 1e-3) on fixed batches of 512/512/176."""
    torch.set_num_threads(1)
    g = golden("mlp_training")
    net = fresh_net()
    tr = oracle_trainer(net)
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


def test_trainer_has_no_cpu_fallback():
    """Without a CUDA device the product's trainer refuses to construct (its only backend is the sm_100a kernels)."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pygent_b200.lib import PygentB200Error
    with pytest.raises(PygentB200Error):
        DynamicsTrainer(fresh_net(), dyn_lr=1e-3, weight_decay=1e-3)
    with pytest.raises(PygentB200Error):
        DynamicsTrainer(fresh_net(), dyn_lr=1e-3, weight_decay=1e-3, device="cpu")


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
        tr = oracle_trainer(net)
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
    single = run_fixed_batches(oracle_trainer(net), dataset(), epochs=1)
    for rank, losses, W3, w2sum, ncoll in res:
        assert ncoll == 3  # one collective per optimiser step
        assert np.allclose(losses, single, rtol=1e-5) and np.allclose(losses, g["losses"][:3], rtol=2e-5)
        assert np.allclose(W3, net.layer3.weight.detach().numpy(), rtol=1e-4, atol=1e-6)
        assert abs(w2sum - float(net.layer2.weight.detach().double().sum())) < 1e-3
    assert np.array_equal(res[0][2], res[1][2]) and res[0][3] == res[1][3]  # ranks stay bit-identical


def test_run_learning_schedule_matches_reference():
    """The outer loop's schedule -- warm-up episodes until D_rand holds max(batch_size, warm_up) samples, the first training,
    then `train_dynamics` every `aggregation_interval`-th step (the reference's `steps % interval == 0 & flag`, i.e.
    `steps % interval == (0 & flag)`), the planner's `save()` unless MPC is used, a checkpoint whenever `episode` is a multiple
    of `checkInterval` -- as the call sequences of the UNMODIFIED `MBRL.run_learning` on a recording stand-in
    (oracle/make_golden_mbrl.py -> tests/golden/mbrl_schedule.json; plotting calls are out of scope and ignored)."""
    import json
    import os
    import types
    from pygent_b200.algorithms.mbrl import MBRL
    with open(os.path.join(os.path.dirname(__file__), "golden", "mbrl_schedule.json")) as fh:
        cases = json.load(fh)
    assert len(cases) == 6
    for case in cases:
        kw, calls = case["kw"], []
        stub = types.SimpleNamespace(batch_size=kw["batch_size"], warm_up=kw["warm_up"], dynamics_first_trained=False,
                                     use_mpc=kw["use_mpc"], aggregation_interval=kw["aggregation_interval"],
                                     checkInterval=kw["check_interval"], episode=1, D_rand=types.SimpleNamespace(data=[]),
                                     D_RL=types.SimpleNamespace(data=[]))

        def episode(kind, stub=stub, calls=calls):
            def f():
                calls.append(kind)
                if kind == "random":
                    stub.D_rand.data.extend([0] * 60)
                stub.episode += 1
            return f
        stub.run_random_episode, stub.run_episode = episode("random"), episode("episode")
        stub.train_dynamics = lambda calls=calls: calls.append("train")
        stub.agent = types.SimpleNamespace(traj_optimizer=types.SimpleNamespace(save=lambda calls=calls: calls.append("ilqr_save"), saving=True))
        stub.save = lambda calls=calls: calls.append("save")
        MBRL.run_learning(stub, kw["n"])
        want = [c for c in case["calls"] if c not in ("plot", "curve", "animation")]
        assert calls == want, (kw, calls, want)
