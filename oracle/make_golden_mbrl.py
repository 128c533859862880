"""Golden vectors of the MBRL outer loop's deterministic pieces from the UNMODIFIED reference:
`MBRL.run_random_episode` (pygent/algorithms/mbrl.py:211-249: random actions through `MPCAgent.take_random_action`,
nmpc.py:279-296, `CartPole.step`, the noisy transition record, `DataSet.force_add_sample`, the episode statistics) and
`MBRL.pred_loss` (:189-209), each called UNBOUND on a stand-in `self` that carries exactly the attributes the method
reads (the reference's own `MBRL.__init__` passes a float layer width to torch and no longer constructs).  The global
numpy generator is seeded, so an implementation that draws in the reference's order -- action, then the noise of x_, u, x,
o_, o -- reproduces every number.     python -m oracle.make_golden_mbrl"""
import os
import types

import numpy as np
import torch

from . import mlp_oracle as mo
from . import ref_harness as rh
from .make_golden import GOLDEN

SEED, T_EPISODE, DT, DATA_NOISE = 123, 1.2, 0.02, 1e-3


def c_k(x, u, mod):  # examples/mbrl/cart_pole_hpc.py:24-28
    x1, x2, x3, x4 = x
    u1, = u
    return 10 * x1 ** 2 + 5 * x2 ** 2 + .01 * x3 ** 2 + .01 * x4 ** 2 + 0.1 * u1 ** 2


def main():
    pg = rh.load_reference()
    import pygent.nn_models as nnm
    from pygent.algorithms.mbrl import MBRL
    from pygent.algorithms.nmpc import MPCAgent
    from pygent.data import DataSet
    torch.set_num_threads(1)
    with rh.quiet():
        env = pg.environments.CartPole(c_k, [0, np.pi, 0, 0], DT)
    agent = types.SimpleNamespace(uMax=env.uMax, uDim=env.uDim, history=np.zeros([1, env.uDim]), tt=[0], u=None)
    agent.take_random_action = lambda dt, ou_noise=True: MPCAgent.take_random_action(agent, dt, ou_noise=ou_noise)

    def agent_reset():
        agent.history, agent.tt = np.zeros([1, env.uDim]), [0]
    agent.reset = agent_reset
    stub = types.SimpleNamespace(environment=env, agent=agent, t=T_EPISODE, dt=DT, data_noise=DATA_NOISE, episode=1,   # Algorithm.__init__ (core.py:26)
                                 D_rand=DataSet(10 ** 6), meanCost=[], totalCost=[], episode_steps=[])
    np.random.seed(SEED)
    with rh.quiet():
        MBRL.run_random_episode(stub)
        MBRL.run_random_episode(stub)   # a second episode: reset semantics, data set growth
    data = stub.D_rand.data
    col = lambda k: np.array([np.asarray(s[k], dtype=float).ravel() for s in data])
    out = {k: col(k) for k in ("x_", "u", "x", "o_", "o", "c")}
    out["terminated"] = col("t")
    out.update(meanCost=np.array(stub.meanCost), totalCost=np.array(stub.totalCost), episode_steps=np.array(stub.episode_steps),
               episode=stub.episode, seed=SEED, t_episode=T_EPISODE, dt=DT, data_noise=DATA_NOISE, x0=np.array([0, np.pi, 0, 0.]))

    # pred_loss of the first 16 transitions under a fixed network and the data set's moments
    w = mo.make_weights(6, 2)
    net = nnm.NNDynamics(4, 1, dxDim=2, oDim=5, xIsAngle=[False, True, False, False])
    with torch.no_grad():
        for i, l in enumerate((net.layer1, net.layer2, net.layer3), 1):
            l.weight.copy_(torch.from_numpy(w["W%d" % i]))
            l.bias.copy_(torch.from_numpy(w["b%d" % i]))
    stub2 = types.SimpleNamespace(nn_dynamics=net, sparse_dyn=True)
    MBRL.set_moments(stub2, stub.D_rand)
    net.eval()
    out["pred_loss"] = np.array([float(MBRL.pred_loss(stub2, s)) for s in data[:16]])
    np.savez_compressed(os.path.join(GOLDEN, "mbrl_warmup.npz"), **out)
    print("mbrl_warmup.npz: %d transitions, episode steps %s, mean costs %s, pred_loss[0] %.6f" % (
        len(data), stub.episode_steps, np.round(stub.meanCost, 6), out["pred_loss"][0]))


def schedule_trace(MBRL, n, aggregation_interval, use_mpc, warm_up, batch_size, check_interval, samples_per_episode=60):
    """The call sequence of `MBRL.run_learning(n)` (mbrl.py:275-310) on a recording stand-in `self`: which of
    run_random_episode / train_dynamics / run_episode / traj_optimizer.save / learning_curve / save / plot it calls, in order."""
    calls = []
    stub = types.SimpleNamespace(batch_size=batch_size, warm_up=warm_up, dynamics_first_trained=False, use_mpc=use_mpc,
                                 aggregation_interval=aggregation_interval, checkInterval=check_interval, episode=1,
                                 D_rand=types.SimpleNamespace(data=[]), D_RL=types.SimpleNamespace(data=[]))

    def episode(kind):
        def f():
            calls.append(kind)
            if kind == "random":
                stub.D_rand.data.extend([0] * samples_per_episode)
            stub.episode += 1
        return f
    stub.run_random_episode, stub.run_episode = episode("random"), episode("episode")
    stub.train_dynamics = lambda: calls.append("train")
    stub.agent = types.SimpleNamespace(traj_optimizer=types.SimpleNamespace(save=lambda: calls.append("ilqr_save"), saving=True))
    stub.learning_curve = lambda: calls.append("curve")
    stub.save = lambda: calls.append("save")
    stub.plot = lambda: calls.append("plot")
    stub.animation = lambda: calls.append("animation")
    with rh.quiet():
        MBRL.run_learning(stub, n)
    return calls


def main_schedule():
    rh.load_reference()
    from pygent.algorithms.mbrl import MBRL
    import json
    cases = []
    for agg in (1, 2, 3):
        for use_mpc in (False, True):
            kw = dict(n=12, aggregation_interval=agg, use_mpc=use_mpc, warm_up=150, batch_size=128, check_interval=4)
            cases.append({"kw": kw, "calls": schedule_trace(MBRL, **kw)})
    with open(os.path.join(GOLDEN, "mbrl_schedule.json"), "w") as fh:
        json.dump(cases, fh, indent=0)
    print("mbrl_schedule.json:", len(cases), "cases;", cases[2]["calls"][:14])


if __name__ == "__main__":
    main()
    main_schedule()
