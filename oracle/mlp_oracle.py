"""CPU restatement of the learned-dynamics path (TEST INFRASTRUCTURE, NOT PRODUCT).

Follows the reference line by line in numpy with the reference's dtypes:
  NNDynamics.ode      pygent/nn_models.py:193-202   (observation in fp64 -> float32, standardise, MLP fp32)
  MBRL.ode            pygent/algorithms/mbrl.py:125-130 (sparse_dyn: [x[n/2:], 1/dt * nn.ode(x, u)], the
                      product float * float32-array stays float32)
  fast_step           pygent/environments.py:313-315 (x + dt * ode, fp64)
Pinned by tests/golden/mlp_cart_pole.npz (oracle/make_golden_mlp.py runs the reference's own classes).
"""
import numpy as np


def make_weights(n_in, dx_dim, seed=0, hidden=500):
    """Deterministic network (same law as torch.nn.Linear's default init, from numpy so that the golden
    generator and the tests build identical weights without shipping them)."""
    rng = np.random.RandomState(seed)
    lin = lambda fo, fi: (rng.uniform(-1, 1, (fo, fi)).astype(np.float32) / np.float32(np.sqrt(fi)),
                          rng.uniform(-1, 1, fo).astype(np.float32) / np.float32(np.sqrt(fi)))
    W1, b1 = lin(hidden, n_in)
    W2, b2 = lin(hidden, hidden)
    W3, b3 = lin(dx_dim, hidden)
    return dict(W1=W1, b1=b1, W2=W2, b2=b2, W3=W3, b3=b3)


def make_moments(n_obs, m, dx_dim, seed=1):
    rng = np.random.RandomState(seed)
    f = lambda k, lo, hi: rng.uniform(lo, hi, k).astype(np.float32)
    return dict(o_mean=f(n_obs, -.2, .2), o_var=f(n_obs, .5, 2.), u_mean=f(m, -.5, .5), u_var=f(m, 2., 6.),
                dx_mean=f(dx_dim, -.01, .01), dx_var=f(dx_dim, .02, .2))


def observation(x, is_angle):
    cols = []
    for i, a in enumerate(is_angle):
        if a:
            cols += [np.cos(x[..., i]), np.sin(x[..., i])]
        else:
            cols.append(x[..., i])
    return np.stack(cols, axis=-1)


def nn_ode(w, mom, is_angle, x, u):
    """NNDynamics.ode for a batch: x [R, n] fp64, u [R, m] -> [R, dx] float32."""
    o = observation(np.asarray(x, dtype=np.float64), is_angle).astype(np.float32)
    o = (o - mom["o_mean"]) / mom["o_var"]
    uu = (np.asarray(u, dtype=np.float64).astype(np.float32) - mom["u_mean"]) / mom["u_var"]
    h = np.concatenate([o, uu], axis=1)
    h1 = np.maximum(h @ w["W1"].T + w["b1"], 0)
    h2 = np.maximum(h1 @ w["W2"].T + w["b2"], 0)
    y = h2 @ w["W3"].T + w["b3"]
    return (y * mom["dx_var"] + mom["dx_mean"]).astype(np.float32)


def mbrl_ode(w, mom, is_angle, x, u, dt):
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[1]
    dv = np.float32(1.0 / dt) * nn_ode(w, mom, is_angle, x, u)
    return np.concatenate([x[:, n // 2:], dv.astype(np.float64)], axis=1)


def rollout(w, mom, is_angle, x0, dt, U=None, gains=None, alpha=1.0, ulim=None, T=None):
    """Euler rollout, x0 [B, n]; U [B, T, m] or gains (KK [B,T,m,n], kk [B,T,m], xbar [B,T+1,n], ubar [B,T,m])."""
    x = np.array(x0, dtype=np.float64)
    B, n = x.shape
    if U is not None:
        T = U.shape[1]
    if gains is not None:
        T = gains[3].shape[1]
    m = mom["u_mean"].shape[0]
    X, Us = [x.copy()], []
    for k in range(T):
        if gains is not None:
            KK, kk, xbar, ubar = gains
            u = np.einsum("bmn,bn->bm", KK[:, k], x - xbar[:, k]) + alpha * kk[:, k] + ubar[:, k]
        elif U is not None:
            u = U[:, k]
        else:
            u = np.zeros((B, m))
        if ulim is not None:
            u = np.clip(u, -np.asarray(ulim), np.asarray(ulim))
        x = x + dt * mbrl_ode(w, mom, is_angle, x, u, dt)
        X.append(x.copy())
        Us.append(u.copy())
    return np.stack(X, axis=1), np.stack(Us, axis=1)
