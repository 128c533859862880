"""Learned-MLP dynamics on the GPU (include/pygent_b200_mlp.h) against the reference's own classes
(tests/golden/mlp_cart_pole.npz, made by oracle/make_golden_mlp.py) and the numpy restatement oracle/mlp_oracle.py."""
import numpy as np
import pytest
import torch

from oracle import mlp_oracle as mo
from pygent_b200 import mlp

pytestmark = pytest.mark.gpu

IS_ANGLE = {4: [False, True, False, False], 6: [False, True, True, False, False, False], 2: [True, False],
            8: [False, True, True, True, False, False, False, False]}   # pendulum, cart-pole/acrobot, double, triple (examples/mbrl)


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1.0)))


def model(n=4, m=1, seed=0):
    ia = IS_ANGLE[n]
    n_obs, dx = n + sum(ia), n // 2
    w, mom = mo.make_weights(n_obs + m, dx, seed=seed), mo.make_moments(n_obs, m, dx, seed=seed + 1)
    dyn = mlp.MLPDynamics.from_arrays(n, m, ia, w["W1"], w["b1"], w["W2"], w["b2"], w["W3"], w["b3"], **mom)
    return dyn, w, mom, ia


# stated tolerances vs the reference's fp32 network, per evaluation: the fp32 FFMA path differs by summation order; the
# reference-precision tensor-core path ("f16x3": fp16 hi/lo split operands, Ah Wh + Al Wh + Ah Wl, fp32 accumulation) is
# stated at 1e-5 (observed ~1e-6); the bf16 tensor-core path rounds H1 and W2 to 8 mantissa bits before the 500-term dot
# products and is an explicitly lower-precision extra
TOL = {"fp32": 2e-6, "f16x3": 1e-5, "bf16": 2e-2}


@pytest.mark.parametrize("precision", ["fp32", "f16x3", "bf16"])
def test_mlp_ode_matches_reference(golden, precision):
    g = golden("mlp_cart_pole")
    dyn, w, mom, ia = model()
    rhs = dyn.ode(g["xs"], g["us"], float(g["dt"]), precision=precision)
    assert rel(rhs, g["rhs"]) < TOL[precision]
    assert np.array_equal(rhs[:, :2], g["xs"][:, 2:])  # the position half is a copy of the velocities


@pytest.mark.parametrize("precision", ["fp32", "f16x3", "bf16"])
def test_mlp_euler_rollout_matches_reference(golden, precision):
    g = golden("mlp_cart_pole")
    dyn, w, mom, ia = model()
    x0 = torch.as_tensor(g["x0"][:, None].copy(), device="cuda")
    U = torch.as_tensor(g["U"][:, :, None].copy(), device="cuda")
    X = dyn.rollout_device(x0, float(g["dt"]), U=U, precision=precision).cpu().numpy()[:, :, 0]
    assert rel(X, g["X"]) < TOL[precision] * 10


@pytest.mark.parametrize("precision", ["fp32", "f16x3", "bf16"])
@pytest.mark.parametrize("n,B,T", [(4, 200, 25), (6, 129, 12), (2, 64, 10), (4, 1, 5), (8, 70, 8)])
def test_mlp_rollouts_match_oracle(precision, n, B, T):
    """Open-loop and closed-loop (iLQR forward-pass law) rollouts, ragged batches, the four state dimensions of examples/mbrl/*.py."""
    dyn, w, mom, ia = model(n=n, seed=n)
    rng = np.random.default_rng(n)
    x0 = rng.uniform(-1, 1, (B, n))
    U = rng.uniform(-6, 6, (B, T, 1))
    dt = 0.02
    Xo, _ = mo.rollout(w, mom, ia, x0, dt, U=U)
    X = dyn.rollout_device(mlp._dev(x0, "cuda"), dt, U=torch.as_tensor(np.ascontiguousarray(U.transpose(1, 2, 0)), device="cuda"),
                           precision=precision)
    assert rel(np.moveaxis(X.cpu().numpy(), -1, 0), Xo) < TOL[precision] * 10
    # closed loop around the open-loop trajectory with random gains, clipped controls
    KK = rng.uniform(-1, 1, (B, T, 1, n)) * 0.5
    kk = rng.uniform(-1, 1, (B, T, 1))
    x0b = x0 + rng.uniform(-0.05, 0.05, (B, n))
    Xc, Uc = mo.rollout(w, mom, ia, x0b, dt, gains=(KK, kk, Xo, U), alpha=0.3, ulim=[5.0])
    d = lambda a: torch.as_tensor(np.ascontiguousarray(np.moveaxis(a, 0, -1)), device="cuda")
    X2, U2 = dyn.rollout_device(mlp._dev(x0b, "cuda"), dt, gains=(d(KK), d(kk), d(Xo), d(U)), alpha=0.3, ulim=[5.0],
                                precision=precision, want_u=True)
    assert rel(np.moveaxis(X2.cpu().numpy(), -1, 0), Xc) < TOL[precision] * 10
    assert rel(np.moveaxis(U2.cpu().numpy(), -1, 0), Uc) < TOL[precision] * 10
    assert float(U2.abs().max()) <= 5.0


def test_mlp_tensor_core_path_tracks_fp32_at_scale():
    """32,768 rows (BASELINE.json config 5): the tensor-core paths vs the fp32 path, per-evaluation error."""
    dyn, w, mom, ia = model()
    rng = np.random.default_rng(0)
    R = 32768
    x = torch.as_tensor(rng.uniform(-2, 2, (4, R)), device="cuda")
    u = torch.as_tensor(rng.uniform(-8, 8, (1, R)), device="cuda")
    a = dyn.ode_device(x, u, 0.02, "fp32").cpu().numpy()
    b = dyn.ode_device(x, u, 0.02, "bf16").cpu().numpy()
    assert rel(b, a) < TOL["bf16"]
    assert np.sqrt(np.mean((a[2:] - b[2:]) ** 2)) / np.sqrt(np.mean(a[2:] ** 2)) < 5e-3
    c = dyn.ode_device(x, u, 0.02, "f16x3").cpu().numpy()
    assert rel(c, a) < TOL["f16x3"]
    assert np.sqrt(np.mean((a[2:] - c[2:]) ** 2)) / np.sqrt(np.mean(a[2:] ** 2)) < 2e-6
    # and against the float64 evaluation of the same network: the split-fp16 path is as close to it as the fp32 path is
    ref = mo.mbrl_ode(w, mom, ia, x.cpu().numpy().T[:2048], u.cpu().numpy().T[:2048], 0.02).T
    assert rel(c[:, :2048], ref) < TOL["f16x3"]


def test_mlp_reference_precision_jacobian_on_tensor_cores():
    """PG_MLP_LIN_ANALYTIC at PG_MLP_F16X3 against the float64 chain rule: with fp32-class operands the ReLU masks agree
    with the exact ones except where a pre-activation is within rounding of zero, so the Jacobian is the exact
    derivative of the piecewise-linear network: entries within 2e-5 of the scale on all but a handful of points
    (a switched unit moves an entry by ~1/sqrt(active units)), Frobenius error < 1e-3."""
    dyn, w, mom, ia = model()
    rng = np.random.default_rng(4)
    T, B, dt = 5, 203, 0.02  # 1015 rows: ragged last tile
    xb = np.array([0, np.pi, 0, 0]) + rng.uniform(-1.5, 1.5, (B, T + 1, 4))
    ub = rng.uniform(-8, 8, (B, T, 1))
    d = lambda a: torch.as_tensor(np.ascontiguousarray(np.moveaxis(a, 0, -1)), device="cuda")
    A, Bm = dyn.linearize_device(d(xb), d(ub), dt, mode="analytic", precision="f16x3")
    Ae, Be = mo.exact_jacobian(w, mom, ia, xb[:, :-1].reshape(-1, 4), ub.reshape(-1, 1), dt)
    A, Bm = np.moveaxis(A.cpu().numpy(), -1, 0).reshape(-1, 4, 4), np.moveaxis(Bm.cpu().numpy(), -1, 0).reshape(-1, 4, 1)
    for g, e in ((A[:, 2:], Ae[:, 2:]), (Bm[:, 2:], Be[:, 2:])):
        err = np.abs(g - e) / np.abs(e).max()
        per_point = err.reshape(len(err), -1).max(axis=1)
        assert np.median(err) < 2e-6 and np.mean(per_point > 2e-5) < 0.01, (np.median(err), np.mean(per_point > 2e-5), err.max())
        assert np.linalg.norm(g - e) < 1e-3 * np.linalg.norm(e)
    assert np.array_equal(A[:, :2], Ae[:, :2]) and np.array_equal(Bm[:, :2], Be[:, :2])


# ---- iLQR over the learned dynamics (BASELINE.json config 5) --------------------------------------------------------
def c_k(x, u, mod):
    x1, x2, x3, x4 = x
    u1, = u
    return 10 * x1 ** 2 + 5 * x2 ** 2 + .01 * x3 ** 2 + .01 * x4 ** 2 + 0.1 * u1 ** 2


def c_N(x, mod):
    x1, x2, x3, x4 = x
    return 100 * x1 ** 2 + 100 * x2 ** 2 + 10 * x3 ** 2 + 10 * x4 ** 2


COST = mo.QuadCost([10, 5, .01, .01], [0.1], [100, 100, 10, 10])


def learned_env(x0, precision, linearization=None, dt=0.02):
    from pygent_b200.environments import LearnedStateSpaceModel
    from pygent_b200.nn_models import NNDynamics
    w, mom = mo.make_weights(6, 2), mo.make_moments(5, 1, 2)
    net = NNDynamics(4, 1, dxDim=2, oDim=5, xIsAngle=IS_ANGLE[4])
    with torch.no_grad():
        for i, l in enumerate((net.layer1, net.layer2, net.layer3), 1):
            l.weight.copy_(torch.from_numpy(w["W%d" % i]))
            l.bias.copy_(torch.from_numpy(w["b%d" % i]))
    for k, a in (("oMean", "o_mean"), ("oVar", "o_var"), ("uMean", "u_mean"), ("uVar", "u_var"), ("dxMean", "dx_mean"), ("dxVar", "dx_var")):
        setattr(net, k, torch.from_numpy(mom[a])[None])
    env = LearnedStateSpaceModel(net, c_k, x0, 1, dt, precision=precision, linearization=linearization)
    env.uMax = np.array([8.0])
    return env, w, mom


def test_mlp_linearize_fd_matches_oracle():
    """pg_mlp_linearize, PG_MLP_LIN_FD on the fp32 path = helpers.system_linearization on MBRL.ode; tolerance is the
    noise of differencing an fp32 network at eps = 1e-3."""
    dyn, w, mom, ia = model()
    rng = np.random.default_rng(3)
    T, B, dt = 7, 37, 0.02
    xb = np.array([0, np.pi, 0, 0]) + rng.uniform(-1, 1, (B, T + 1, 4))
    ub = rng.uniform(-6, 6, (B, T, 1))
    d = lambda a: torch.as_tensor(np.ascontiguousarray(np.moveaxis(a, 0, -1)), device="cuda")
    A, Bm = dyn.linearize_device(d(xb), d(ub), dt, mode="fd", precision="fp32")
    Ao, Bo = mo.fd_linearize(w, mom, ia, xb[:, :-1].reshape(-1, 4), ub.reshape(-1, 1), dt)
    A, Bm = np.moveaxis(A.cpu().numpy(), -1, 0).reshape(-1, 4, 4), np.moveaxis(Bm.cpu().numpy(), -1, 0).reshape(-1, 4, 1)
    assert rel(A, Ao) < 2e-3 and rel(Bm, Bo) < 2e-3
    assert np.max(np.abs(A[:, :2, :2])) < 1e-9 and np.max(np.abs(A[:, 0, 2] - 1)) < 1e-9  # position rows: [0 I]


def test_mlp_linearize_analytic_on_tensor_cores_matches_exact_jacobian():
    """PG_MLP_LIN_ANALYTIC (three tcgen05 GEMM passes) against the float64 chain rule.  The Jacobian of a ReLU network
    is piecewise constant: with bf16 operands a few near-zero units switch side, and each switch moves an entry by
    ~1/sqrt(active units) of its scale.  So the bulk of the entries agrees to bf16 rounding (median error 4e-4 of the
    scale measured, bound 2e-3) while ~8% of them carry a switch (Frobenius error 5e-2 measured, bound 8e-2) -- the
    reference's own central differences through the fp32 network sit at 2e-2 for the same reason (kinks within eps)."""
    dyn, w, mom, ia = model()
    rng = np.random.default_rng(4)
    T, B, dt = 5, 203, 0.02  # 1015 rows: ragged last tile
    xb = np.array([0, np.pi, 0, 0]) + rng.uniform(-1.5, 1.5, (B, T + 1, 4))
    ub = rng.uniform(-8, 8, (B, T, 1))
    d = lambda a: torch.as_tensor(np.ascontiguousarray(np.moveaxis(a, 0, -1)), device="cuda")
    A, Bm = dyn.linearize_device(d(xb), d(ub), dt, mode="analytic", precision="bf16")
    Ae, Be = mo.exact_jacobian(w, mom, ia, xb[:, :-1].reshape(-1, 4), ub.reshape(-1, 1), dt)
    A, Bm = np.moveaxis(A.cpu().numpy(), -1, 0).reshape(-1, 4, 4), np.moveaxis(Bm.cpu().numpy(), -1, 0).reshape(-1, 4, 1)
    for g, e in ((A[:, 2:], Ae[:, 2:]), (Bm[:, 2:], Be[:, 2:])):
        err = np.abs(g - e) / np.abs(e).max()
        assert np.median(err) < 2e-3 and err.max() < 0.5
        assert np.linalg.norm(g - e) < 8e-2 * np.linalg.norm(e)
    assert np.array_equal(A[:, :2], Ae[:, :2]) and np.array_equal(Bm[:, :2], Be[:, :2])
    # same points through the reference's scheme on the fp32 path: the two approximations of the same derivative
    Af, Bf = dyn.linearize_device(d(xb), d(ub), dt, mode="fd", precision="fp32")
    Af = np.moveaxis(Af.cpu().numpy(), -1, 0).reshape(-1, 4, 4)
    assert np.linalg.norm(Af[:, 2:] - Ae[:, 2:]) < 5e-2 * np.linalg.norm(Ae[:, 2:])


@pytest.mark.parametrize("precision", ["fp32", "f16x3", "bf16"])
def test_mlp_forward_all_alphas_equals_single_rollouts(precision):
    """pg_mlp_forward: every (alpha, trajectory) row equals the one-alpha rollout bit for bit; a compacted work list
    touches exactly the listed trajectories."""
    dyn, w, mom, ia = model()
    rng = np.random.default_rng(5)
    B, T, dt = 150, 9, 0.02
    d = lambda a: torch.as_tensor(np.ascontiguousarray(np.moveaxis(a, 0, -1)), device="cuda")
    x0 = d(rng.uniform(-1, 1, (B, 4)))
    gains = (d(rng.uniform(-.5, .5, (B, T, 1, 4))), d(rng.uniform(-1, 1, (B, T, 1))), d(rng.uniform(-1, 1, (B, T + 1, 4))),
             d(rng.uniform(-3, 3, (B, T, 1))))
    alphas = [1.0, 0.5, 0.1]
    X, U = dyn.forward_device(x0, dt, alphas, gains, ulim=[5.0], precision=precision)
    for i, al in enumerate(alphas):
        X1, U1 = dyn.rollout_device(x0, dt, gains=gains, alpha=al, ulim=[5.0], precision=precision, want_u=True)
        assert torch.equal(X[i], X1) and torch.equal(U[i], U1)
    idx = torch.tensor(sorted(rng.choice(B, 40, replace=False)) + [0] * (B - 40), dtype=torch.int32, device="cuda")
    cnt = torch.tensor([40], dtype=torch.int32, device="cuda")
    out = {"X": torch.full_like(X, -7.0), "U": torch.full_like(U, -7.0)}
    dyn.forward_device(x0, dt, alphas, gains, ulim=[5.0], precision=precision, out=out, work=(idx, cnt))
    sel = idx[:40].long()
    mask = torch.zeros(B, dtype=torch.bool, device="cuda")
    mask[sel] = True
    assert torch.equal(out["X"][..., sel], X[..., sel]) and torch.equal(out["U"][..., sel], U[..., sel])
    assert bool((out["X"][..., ~mask] == -7.0).all())


@pytest.mark.parametrize("B,T,n_alpha", [(150, 9, 3), (2048, 13, 11), (1900, 100, 11), (9600, 7, 2), (19000, 3, 1)])
def test_mlp_balanced_schedule_is_bit_identical(B, T, n_alpha):
    """pg_mlp_forward_balanced (one cluster per SM pair, equal shares of the (tile, interval) sequence, tiles handed
    over between clusters through X) against pg_mlp_forward (one cluster per tile pair): same bits, with fewer tile
    pairs than SM pairs, with more (every cluster's range then starts and ends inside a tile), and with a work list."""
    dyn, w, mom, ia = model()
    rng = np.random.default_rng(11)
    dt = 0.02
    d = lambda a: torch.as_tensor(np.ascontiguousarray(np.moveaxis(a, 0, -1)), device="cuda")
    x0 = d(rng.uniform(-1, 1, (B, 4)))
    gains = (d(rng.uniform(-.5, .5, (B, T, 1, 4))), d(rng.uniform(-1, 1, (B, T, 1))), d(rng.uniform(-1, 1, (B, T + 1, 4))),
             d(rng.uniform(-3, 3, (B, T, 1))))
    alphas = list(10 ** np.linspace(0, -3, 11))[:n_alpha]
    assert dyn.balanced
    _check_balanced(dyn, x0, dt, alphas, gains, d, rng, B, T, "f16x3")
    _check_balanced(dyn, x0, dt, alphas, gains, d, rng, B, T, "bf16")


def _check_balanced(dyn, x0, dt, alphas, gains, d, rng, B, T, prec):
    Xb, Ub = dyn.forward_device(x0, dt, alphas, gains, ulim=[5.0], precision=prec)
    dyn.balanced = False
    Xp, Up = dyn.forward_device(x0, dt, alphas, gains, ulim=[5.0], precision=prec)
    dyn.balanced = True
    assert torch.equal(Xb, Xp) and torch.equal(Ub, Up)
    # open loop through rollout_device, and a compacted work list
    U = d(rng.uniform(-3, 3, (B, T, 1)))
    Xo = dyn.rollout_device(x0, dt, U=U, precision=prec)
    dyn.balanced = False
    assert torch.equal(Xo, dyn.rollout_device(x0, dt, U=U, precision=prec))
    dyn.balanced = True
    k = max(1, (B * 2) // 3)
    idx = torch.tensor(sorted(rng.choice(B, k, replace=False)) + [0] * (B - k), dtype=torch.int32, device="cuda")
    cnt = torch.tensor([k], dtype=torch.int32, device="cuda")
    out = {"X": torch.full_like(Xb, -7.0), "U": torch.full_like(Ub, -7.0)}
    dyn.forward_device(x0, dt, alphas, gains, ulim=[5.0], precision=prec, out=out, work=(idx, cnt))
    sel = idx[:k].long()
    mask = torch.zeros(B, dtype=torch.bool, device="cuda")
    mask[sel] = True
    assert torch.equal(out["X"][..., sel], Xb[..., sel]) and torch.equal(out["U"][..., sel], Ub[..., sel])
    assert bool((out["X"][..., ~mask] == -7.0).all())


def test_backward_ext_and_traj_cost_match_oracle():
    """pg_ilqr_backward_ext on given Jacobians = ilqr.py:233-346 restated in numpy; pg_traj_cost = the summed
    fast_step costs + final cost."""
    env, w, mom = learned_env(np.zeros((3, 4)), "fp32")
    from pygent_b200.algorithms.ilqr import iLQR
    rng = np.random.default_rng(6)
    B, T, dt = 3, 25, 0.02
    x0 = np.array([0.2, np.pi - .3, .1, -.2]) + rng.uniform(-.1, .1, (B, 4))
    env.reset(x0)
    alg = iLQR(env, T * dt, dt, fastForward=True, fcost=c_N, printing=False, final_backpass=False)
    U = rng.uniform(-3, 3, (B, T, 1))
    X, _ = mo.rollout(w, mom, IS_ANGLE[4], x0, dt, U=U)
    d = lambda a: torch.as_tensor(np.ascontiguousarray(np.moveaxis(a, 0, -1)), device="cuda")
    c = alg.lib.traj_cost(d(X)[None], d(U)[None], dt=dt)
    co = [mo.traj_cost(COST, X[b], U[b], dt) for b in range(B)]
    assert rel(c["cost"][0].cpu().numpy(), co) < 1e-12
    A = rng.uniform(-1, 1, (B, T, 4, 4))
    Bm = rng.uniform(-1, 1, (B, T, 4, 1))
    mu = torch.full((B,), 1e-6, dtype=torch.float64, device="cuda")
    r = alg.lib.ilqr_backward_ext(d(X), d(U), mu, d(A), d(Bm), dt=dt)
    for b in range(B):
        V1, V2, ok, KK, kk = mo.ilqr_backward(COST, A[b], Bm[b], X[b], U[b], 1e-6, dt)
        assert ok and bool(r["success"][b])
        assert rel(r["KK"][..., b].cpu().numpy(), np.array(KK)) < 1e-8
        assert rel(r["kk"][..., b].cpu().numpy(), np.array(kk).reshape(T, 1)) < 1e-8
        assert rel(r["dV"][:, b].cpu().numpy(), [np.sum(V1), np.sum(V2)]) < 1e-8
    # finite_diff=True (a bare MPCAgent's default, nmpc.py:139): the same sweep with the cost expansion by central
    # differences; this cost is quadratic, so the differences are exact up to their rounding noise (1e-10 / 4 eps^2)
    rf = alg.lib.ilqr_backward_ext(d(X), d(U), mu, d(A), d(Bm), dt=dt, cost_fd=True)
    assert bool(rf["success"].all())
    assert rel(rf["KK"].cpu().numpy(), r["KK"].cpu().numpy()) < 1e-6
    assert rel(rf["kk"].cpu().numpy(), r["kk"].cpu().numpy()) < 1e-6
    assert not torch.equal(rf["KK"], r["KK"])  # it did take the other path
    alg_fd = iLQR(env, T * dt, dt, fastForward=True, fcost=c_N, printing=False, final_backpass=False, finite_diff=True,
                  maxIters=3)
    alg2 = iLQR(env, T * dt, dt, fastForward=True, fcost=c_N, printing=False, final_backpass=False, maxIters=3)
    _, _, c_fd = alg_fd.run_optim()
    _, _, c_an = alg2.run_optim()
    assert rel(c_fd, c_an) < 1e-5


def test_learned_ilqr_matches_reference_planner(golden):
    """The whole learned-dynamics planner on the reference precision path (fp32 network, central differences) against
    the reference's iLQR on StateSpaceModel(MBRL.ode) -- tolerances: see test_mlp_ilqr_oracle_matches_reference."""
    from pygent_b200.algorithms.ilqr import iLQR
    g = golden("mlp_ilqr_cart_pole")
    T, dt = int(g["T"]), float(g["dt"])
    env, w, mom = learned_env(list(g["x0"]), "fp32", "fd")
    alg = iLQR(env, T * dt, dt, fastForward=True, fcost=c_N, constrained=False, printing=False, maxIters=12, final_backpass=False)
    assert rel(alg.cost, g["cost0"]) < 1e-6 and rel(alg.xx, g["xx0"]) < 1e-6
    Ad, Bd, _ = alg.linearization(g["xx0"][:-1], np.zeros((T, 1)))
    assert rel(Ad, g["Ad0"]) < 5e-5 and rel(Bd, g["Bd0"]) < 5e-5
    V1, V2, ok, KK, kk = alg.backward_pass()
    assert ok and rel(KK, g["KK1"]) < 2e-2 and rel(kk[..., 0], g["kk1"]) < 2e-2
    assert rel(V1, g["sV1"]) < 1e-3 and rel(V2, g["sV2"]) < 1e-3
    xx, uu, cost, _ = alg.forward_pass(1.0, KK, kk)
    assert rel(cost, g["cost_alpha1"]) < 1e-5
    xx, uu, cost = alg.run_optim()
    assert rel(cost, g["cost"]) < 1e-4 and rel(xx, g["xx"]) < 5e-3
    assert alg.iterations == 12


def test_learned_ilqr_tensor_core_path_batched():
    """bf16 rollouts + analytic tensor-core Jacobians, a ragged batch with work lists: costs fall monotonically, every
    trajectory ends close to the fp32 / finite-difference planner, and a trajectory's result does not depend on its
    neighbours in the batch."""
    from pygent_b200.algorithms.ilqr import iLQR
    rng = np.random.default_rng(8)
    B, T, dt = 197, 40, 0.02
    x0 = np.array([0.3, np.pi - 0.4, 0.2, -0.5]) + rng.uniform(-.2, .2, (B, 4))
    res = {}
    for prec, lin in (("bf16", "analytic"), ("f16x3", "analytic"), ("fp32", "fd")):
        env, _, _ = learned_env(x0, prec, lin)
        alg = iLQR(env, T * dt, dt, fastForward=True, fcost=c_N, printing=False, maxIters=15, final_backpass=False)
        c0 = alg.cost.copy()
        alg.run_optim()
        res[prec] = (c0, alg.cost.copy(), alg.iters.copy())
        assert np.all(alg.cost <= c0 + 1e-12) and np.all(np.isfinite(alg.cost))
    assert rel(res["bf16"][0], res["fp32"][0]) < 5e-3
    assert np.max(np.abs(res["bf16"][1] - res["fp32"][1]) / res["fp32"][1]) < 1e-2
    assert rel(res["f16x3"][0], res["fp32"][0]) < 1e-6  # the initial rollouts: same network to fp32 rounding
    # same optimum as the reference-precision planner; the two differ in HOW they linearise (exact derivative vs the
    # reference's central differences with eps = 1e-3 through fp32), not in the model
    assert np.max(np.abs(res["f16x3"][1] - res["fp32"][1]) / res["fp32"][1]) < 2e-3
    sub = [5, 77, 196]
    env, _, _ = learned_env(x0[sub], "bf16", "analytic")
    alg = iLQR(env, T * dt, dt, fastForward=True, fcost=c_N, printing=False, maxIters=15, final_backpass=False)
    alg.run_optim()
    assert np.array_equal(alg.cost, res["bf16"][1][sub]) and np.array_equal(alg.iters, res["bf16"][2][sub])


def test_mbrl_outer_loop_runs_end_to_end(tmp_path):
    """MBRL.run_learning on the cart-pole with small settings (reference loop: mbrl.py:283-310): warm-up episodes fill
    D_rand, the model is trained (loss falls), an episode is planned on the learned model with the tensor-core planner and
    executed on the real environment, transitions are gated by the prediction error, and the checkpoint round-trips."""
    from pygent_b200.algorithms.mbrl import MBRL
    from pygent_b200.environments import CartPole

    def cost(x, u, mod):
        x1, x2, x3, x4 = x
        u1, = u
        return 10 * x1 ** 2 + 5 * x2 ** 2 + .01 * x3 ** 2 + .01 * x4 ** 2 + 0.1 * u1 ** 2
    env = CartPole(cost, [0, np.pi, 0, 0], 0.02)
    alg = MBRL(env, 1.0, 0.02, path=str(tmp_path) + "/", warm_up_episodes=12, batch_size=256, training_epochs=8, fcost=c_N,
               maxIters=6, checkInterval=1000, dyn_lr=1e-3)
    while len(alg.D_rand.data) < max(alg.batch_size, alg.warm_up):
        alg.run_random_episode()
    n_rand = len(alg.D_rand.data)
    assert n_rand >= 600 and alg.episode >= 13
    tr = alg.D_rand.data[5]
    assert set(tr) == {"x_", "u", "x", "o_", "o", "c", "t"} and len(tr["o"]) == 5
    losses = alg.train_dynamics()
    assert losses[-1] < 0.5 * losses[0]
    rhs = alg.ode(0.0, np.array([0.1, 3.0, 0.2, -0.1]), np.array([1.0]))
    assert rhs.shape == (4,) and rhs[0] == pytest.approx(0.2) and np.all(np.isfinite(rhs))
    alg.run_learning(1)                          # one planned episode on the real plant
    assert len(alg.environment.history) > 5 and np.all(np.isfinite(alg.environment.history))
    assert np.all(np.abs(alg.agent.history) <= 10.0 + 1e-9)
    assert len(alg.D_RL.data) <= len(alg.environment.history) - 1 and len(alg.optim_time) == 1
    alg.save()
    w = alg.nn_dynamics.layer2.weight.detach().cpu().clone()
    other = MBRL(CartPole(cost, [0, np.pi, 0, 0], 0.02), 1.0, 0.02, path=str(tmp_path) + "/", fcost=c_N, maxIters=6)
    other.load()
    assert torch.equal(other.nn_dynamics.layer2.weight.detach().cpu(), w) and len(other.D_rand.data) == n_rand
    assert other.episode == alg.episode


def test_mbrl_warm_up_episodes_and_pred_loss_match_reference(tmp_path, golden):
    """The deterministic pieces of the MBRL outer loop against the UNMODIFIED reference (oracle/make_golden_mbrl.py):
    two `run_random_episode` calls (mbrl.py:211-249) with the numpy generator seeded like the reference's global one --
    random actions, plant steps (the reference: native RK45; here RK4 x 4), the noisy transition records in the reference's
    draw order (x_, u, x, o_, o), the episode statistics -- and `pred_loss` (:189-209) of the first transitions under a
    fixed network and the data set's moments, through the training kernels' forward pass."""
    from pygent_b200.algorithms.mbrl import MBRL
    from pygent_b200.environments import CartPole
    g = golden("mbrl_warmup")

    def cost(x, u, mod):
        x1, x2, x3, x4 = x
        u1, = u
        return 10 * x1 ** 2 + 5 * x2 ** 2 + .01 * x3 ** 2 + .01 * x4 ** 2 + 0.1 * u1 ** 2
    env = CartPole(cost, list(g["x0"]), float(g["dt"]))
    alg = MBRL(env, float(g["t_episode"]), float(g["dt"]), path=str(tmp_path) + "/", data_noise=float(g["data_noise"]), fcost=c_N, maxIters=4)
    alg.rng = alg.agent.rng = np.random.RandomState(int(g["seed"]))   # ONE stream for actions and noise, as in the reference
    alg.run_random_episode()
    alg.run_random_episode()
    data = alg.D_rand.data
    assert len(data) == len(g["x_"]) and alg.episode == int(g["episode"])
    assert list(alg.episode_steps) == list(g["episode_steps"])
    col = lambda k: np.array([np.asarray(s[k], dtype=float).ravel() for s in data])
    assert np.array_equal(col("t"), g["terminated"])
    assert np.max(np.abs(col("u") - g["u"])) < 1e-12                      # the same draws
    for k in ("x_", "x", "o_", "o"):                                      # noise identical, states to the integrator's tolerance
        assert np.max(np.abs(col(k) - g[k])) < 2e-5, (k, np.max(np.abs(col(k) - g[k])))
    assert rel(col("c"), g["c"]) < 1e-5
    assert rel(np.array(alg.meanCost), g["meanCost"]) < 1e-5 and rel(np.array(alg.totalCost), g["totalCost"]) < 1e-5
    # pred_loss: the golden's network, moments from the data set as MBRL.set_moments computes them
    w = mo.make_weights(6, 2)
    with torch.no_grad():
        for i, l in enumerate((alg.nn_dynamics.layer1, alg.nn_dynamics.layer2, alg.nn_dynamics.layer3), 1):
            l.weight.copy_(torch.from_numpy(w["W%d" % i]).to(l.weight.device))
            l.bias.copy_(torch.from_numpy(w["b%d" % i]).to(l.bias.device))
    alg.nn_dynamics.invalidate()
    alg.trainer.set_moments(alg.D_rand)
    golden_tr = [{k: g[k][i] for k in ("x_", "u", "x", "o_", "o")} for i in range(len(g["pred_loss"]))]
    mine = np.array([alg.pred_loss(tr) for tr in golden_tr])
    assert rel(mine, g["pred_loss"]) < 1e-4, rel(mine, g["pred_loss"])
