"""Batched receding-horizon control (pygent_b200/algorithms/nmpc.py) against closed-loop runs of the unmodified
reference's NMPC/MPCAgent (tests/golden/nmpc_cart_pole.npz, made by oracle/make_golden_nmpc.py: cart-pole of
examples/nmpc/cart_pole.py, horizon 2 s, converged initial plan, 50 control steps, RK4 x 4, constrained=False)."""
import numpy as np
import pytest
import torch

from pygent_b200.algorithms.nmpc import NMPC
from pygent_b200.examples import CASES, make_env

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1.0)))


def controller(x0, horizon, dt, init_optim=True):
    _, c_N, _ = CASES["cart_pole"][3]()
    env, mpc_env = make_env("cart_pole", x0=x0), make_env("cart_pole", x0=x0)
    return NMPC(env, mpc_env, 1.0, dt, horizon, fcost=c_N, constrained=False, maxIters=500, step_iterations=1, ilqr_print=False,
                init_optim=init_optim)


def load_reference_plan(alg, g):
    """The plan the reference holds after its initial solve.  The solve ends, in every run, with a streak of rejected
    line searches followed by an acceptance whose improvement is rounding noise ("red.: 0.00000"): which alpha is
    `current_alpha` afterwards is not reproducible across implementations, yet MPCAgent keeps multiplying it into the
    feed-forward term -- so the closed loops are compared from the SAME plan."""
    alg.agent.load_plan(g["init_xx"], g["init_uu"], g["init_KK"], g["init_kk"][..., None][..., 0], float(g["init_cost"]),
                        float(g["init_alpha"]), env_x=g["init_env_x"])


def test_nmpc_stabilisation_matches_reference(golden):
    """40 control steps around the upright equilibrium (24 accepted plans, 16 rejected line searches): the whole closed
    loop -- plant states, controls, transition costs, the plan's cost bookkeeping and its appended terminal states."""
    g = golden("nmpc_cart_pole_stab")
    dt, horizon, n_steps = float(g["dt"]), float(g["horizon"]), int(g["n_steps"])
    alg = controller(list(g["x0"]), horizon, dt)
    opt = alg.agent.traj_optimizer
    assert rel(opt.cost, g["init_cost"]) < 1e-7 and rel(opt.xx, g["init_xx"]) < 1e-6 and rel(opt.uu, g["init_uu"]) < 1e-5
    assert rel(opt.KK, g["init_KK"]) < 1e-5 and rel(opt.kk[..., 0], g["init_kk"]) < 1e-5
    load_reference_plan(alg, g)
    cost = alg.run(steps=n_steps, printing=False)
    X, U = alg.environment.history, alg.agent.history[1:]
    assert X.shape == g["X"].shape and U.shape == g["U"].shape
    assert rel(X, g["X"]) < 1e-6 and rel(U, g["U"]) < 1e-5 and rel(cost, g["costs"]) < 1e-6
    assert rel(opt.cost, g["plan_costs"][-1]) < 1e-6 and rel(opt.xx[-1], g["plan_xN"][-1]) < 1e-6


def test_nmpc_swing_up_matches_reference(golden):
    """The example of examples/nmpc/cart_pole.py (horizon 2 s, converged initial plan).  In this run the inputs sit at
    +-uMax all the time and the reference's plan tail explodes (its appended terminal states come from REJECTED
    line-search candidates; plan costs reach 1e5 by step 33 and 1e38 by step 40), so rounding differences are amplified
    without bound: the comparison covers the 30 steps in which the plan is still conditioned (plan terminal state
    equal to 4e-5), where the plant follows the reference to 1e-13."""
    g = golden("nmpc_cart_pole")
    dt, horizon, n = float(g["dt"]), float(g["horizon"]), 30
    alg = controller(list(g["x0"]), horizon, dt)
    opt = alg.agent.traj_optimizer
    assert rel(opt.cost, g["init_cost"]) < 1e-7 and rel(opt.xx, g["init_xx"]) < 1e-5 and rel(opt.uu, g["init_uu"]) < 1e-4
    assert rel(alg.agent._xe.cpu().numpy()[:, 0], g["init_env_x"]) < 1e-5  # where the planner's environment was left
    # the final backward pass (equilibrium of the final cost + discrete Riccati equation) replaced the gains
    assert rel(opt.KK, g["init_KK"]) < 1e-5 and rel(opt.kk[..., 0], g["init_kk"]) < 1e-5 and opt.mu == float(g["init_mu"]) == 0.0
    load_reference_plan(alg, g)
    cost = alg.run(steps=n, printing=False)
    X, U = alg.environment.history, alg.agent.history[1:]
    assert rel(X, g["X"][:n + 1]) < 1e-9 and rel(U, g["U"][:n]) < 1e-9 and rel(cost, g["costs"][:n]) < 1e-9
    assert rel(opt.cost, g["plan_costs"][n - 1]) < 1e-6   # 34261.92...: the bookkeeping of the exploding plan still agrees
    assert rel(opt.xx[-1], g["plan_xN"][n - 1]) < 1e-4


def test_nmpc_batch_is_independent_plants(golden):
    """Every plant of a batch behaves exactly like the same plant controlled alone; noise only when asked for."""
    g = golden("nmpc_cart_pole")
    dt, horizon = float(g["dt"]), 1.0
    rng = np.random.default_rng(2)
    x0 = np.array([0, np.pi, 0, 0]) + rng.uniform(-0.2, 0.2, (37, 4))
    x0[0] = g["x0"]
    alg = controller(x0, horizon, dt)
    cost = alg.run(steps=12, printing=False)
    X, U = alg.environment.history, alg.agent.history
    assert X.shape == (37, 13, 4) and U.shape == (37, 13, 1) and cost.shape == (37, 12)
    assert np.all(np.abs(U) <= 10.0 + 1e-12)
    for b in (0, 5, 36):
        one = controller(list(x0[b]), horizon, dt)
        c1 = one.run(steps=12, printing=False)
        assert np.array_equal(one.environment.history, X[b]) and np.array_equal(one.agent.history, U[b])
        assert np.array_equal(c1, cost[b])
    up = np.array([0.05, 0.05, 0.0, 0.0]) + rng.uniform(-0.01, 0.01, (3, 4))  # near the upright: inputs not saturated
    quiet, noisy = controller(up, horizon, dt), controller(up, horizon, dt)
    noisy.agent.add_noise, noisy.agent.noise_gain = True, 0.01
    quiet.run(steps=3, printing=False)
    noisy.run(steps=3, printing=False)
    d = np.abs(noisy.agent.history[:, 1] - quiet.agent.history[:, 1])
    assert np.all(d > 0) and np.all(d < 10 * 0.01 * 6)  # uMax * N(0, noise_gain)


def test_nmpc_cuda_graph_replay_equals_eager_loop():
    """The control step captured into a CUDA graph (possible once the planner's terminal value is cached) replays to
    exactly the eager loop's states, controls and costs."""
    rng = np.random.default_rng(5)
    x0 = np.array([0.1, 0.2, 0.0, 0.0]) + rng.uniform(-0.05, 0.05, (64, 4))
    runs = []
    for use_graph in (False, True):
        _, c_N, _ = CASES["cart_pole"][3]()
        alg = NMPC(make_env("cart_pole", x0=x0), make_env("cart_pole", x0=x0), 1.0, 0.02, 1.0, fcost=c_N, constrained=False,
                   maxIters=50, step_iterations=1, cache_terminal_value=True)
        assert alg._graph_ok()
        cost = alg.run(steps=15, printing=False, use_graph=use_graph)
        runs.append((alg.environment.history.copy(), alg.agent.history.copy(), cost.copy(), alg.agent.traj_optimizer.cost.copy()))
    for a, b in zip(*runs):
        assert np.array_equal(a, b)
    plain = controller(x0[:2], 1.0, 0.02)
    assert not plain._graph_ok()  # final backward pass re-solved on the host every step
    with pytest.raises(ValueError):
        plain.run(steps=6, printing=False, use_graph=True)


@pytest.mark.parametrize("name,fixture", [("double", "nmpc_double_serial"), ("triple", "nmpc_triple")])
def test_nmpc_pendulum_chains_match_reference(golden, name, fixture):
    """examples/nmpc/double.py (CartPoleDoubleSerial, n = 6) and triple.py (CartPoleTriple, n = 8): the cost regulating
    around the hanging configuration, cart displaced by 0.5 m -- converged initial plan with the final backward pass at the
    final cost's equilibrium, then the closed loop (accepted plans and rejected line searches) against the unmodified
    reference's."""
    from pygent_b200.examples import make_nmpc_example
    g = golden(fixture)
    dt, horizon, n_steps = float(g["dt"]), float(g["horizon"]), int(g["n_steps"])
    alg = make_nmpc_example(name, list(g["x0"]), horizon=horizon, maxIters=int(g["init_iterations"]))
    assert alg.dt == dt
    opt = alg.agent.traj_optimizer
    assert rel(opt.cost, g["init_cost"]) < 1e-7 and rel(opt.xx, g["init_xx"]) < 1e-6 and rel(opt.uu, g["init_uu"]) < 1e-5
    assert rel(opt.KK, g["init_KK"]) < 1e-5 and rel(opt.kk[..., 0], g["init_kk"]) < 1e-5
    load_reference_plan(alg, g)
    cost = alg.run(steps=n_steps, printing=False)
    X, U = alg.environment.history, alg.agent.history[1:]
    assert X.shape == g["X"].shape and U.shape == g["U"].shape
    assert rel(X, g["X"]) < 1e-6 and rel(U, g["U"]) < 1e-5 and rel(cost, g["costs"]) < 1e-6
    assert rel(opt.cost, g["plan_costs"][-1]) < 1e-6 and rel(opt.xx[-1], g["plan_xN"][-1]) < 1e-6
