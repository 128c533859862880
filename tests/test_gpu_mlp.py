"""Learned-MLP dynamics on the GPU (include/pygent_b200_mlp.h) against the reference's own classes
(tests/golden/mlp_cart_pole.npz, made by oracle/make_golden_mlp.py) and the numpy restatement oracle/mlp_oracle.py."""
import numpy as np
import pytest
import torch

from oracle import mlp_oracle as mo
from pygent_b200 import mlp

pytestmark = pytest.mark.gpu

IS_ANGLE = {4: [False, True, False, False], 6: [False, True, True, False, False, False], 2: [True, False]}


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1.0)))


def model(n=4, m=1, seed=0):
    ia = IS_ANGLE[n]
    n_obs, dx = n + sum(ia), n // 2
    w, mom = mo.make_weights(n_obs + m, dx, seed=seed), mo.make_moments(n_obs, m, dx, seed=seed + 1)
    dyn = mlp.MLPDynamics.from_arrays(n, m, ia, w["W1"], w["b1"], w["W2"], w["b2"], w["W3"], w["b3"], **mom)
    return dyn, w, mom, ia


# stated tolerances: fp32 path tracks the fp32 reference network to summation order; the tensor-core path rounds
# H1 and W2 to bf16 (8 mantissa bits) before the 500-term dot products
TOL = {"fp32": 2e-6, "bf16": 2e-2}


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_mlp_ode_matches_reference(golden, precision):
    g = golden("mlp_cart_pole")
    dyn, w, mom, ia = model()
    rhs = dyn.ode(g["xs"], g["us"], float(g["dt"]), precision=precision)
    assert rel(rhs, g["rhs"]) < TOL[precision]
    assert np.array_equal(rhs[:, :2], g["xs"][:, 2:])  # the position half is a copy of the velocities


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_mlp_euler_rollout_matches_reference(golden, precision):
    g = golden("mlp_cart_pole")
    dyn, w, mom, ia = model()
    x0 = torch.as_tensor(g["x0"][:, None].copy(), device="cuda")
    U = torch.as_tensor(g["U"][:, :, None].copy(), device="cuda")
    X = dyn.rollout_device(x0, float(g["dt"]), U=U, precision=precision).cpu().numpy()[:, :, 0]
    assert rel(X, g["X"]) < TOL[precision] * 10


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("n,B,T", [(4, 200, 25), (6, 129, 12), (2, 64, 10), (4, 1, 5)])
def test_mlp_rollouts_match_oracle(precision, n, B, T):
    """Open-loop and closed-loop (iLQR forward-pass law) rollouts, ragged batches, three state dimensions."""
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
    """32,768 rows (BASELINE.json config 5): bf16 tensor-core layer 2 vs the fp32 path, per-evaluation error."""
    dyn, w, mom, ia = model()
    rng = np.random.default_rng(0)
    R = 32768
    x = torch.as_tensor(rng.uniform(-2, 2, (4, R)), device="cuda")
    u = torch.as_tensor(rng.uniform(-8, 8, (1, R)), device="cuda")
    a = dyn.ode_device(x, u, 0.02, "fp32").cpu().numpy()
    b = dyn.ode_device(x, u, 0.02, "bf16").cpu().numpy()
    assert rel(b, a) < TOL["bf16"]
    assert np.sqrt(np.mean((a[2:] - b[2:]) ** 2)) / np.sqrt(np.mean(a[2:] ** 2)) < 5e-3
