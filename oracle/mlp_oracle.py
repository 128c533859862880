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


# ---- iLQR over the learned dynamics (the reference's MBRL planner: NMPC/iLQR on StateSpaceModel(MBRL.ode)) --------
def fd_linearize(w, mom, is_angle, x, u, dt, eps=1e-3):
    """helpers.system_linearization (pygent/helpers.py:130-148) applied to MBRL.ode, batched: x [R, n], u [R, m] ->
    A [R, n, n], B [R, n, m] by central differences with eps = 1e-3."""
    x, u = np.asarray(x, dtype=np.float64), np.asarray(u, dtype=np.float64)
    R, n = x.shape
    m = u.shape[1]
    A, B = np.zeros((R, n, n)), np.zeros((R, n, m))
    for i in range(n):
        e = np.zeros(n)
        e[i] = eps
        A[:, :, i] = (mbrl_ode(w, mom, is_angle, x + e, u, dt) - mbrl_ode(w, mom, is_angle, x - e, u, dt)) / (2 * eps)
    for j in range(m):
        e = np.zeros(m)
        e[j] = eps
        B[:, :, j] = (mbrl_ode(w, mom, is_angle, x, u + e, dt) - mbrl_ode(w, mom, is_angle, x, u - e, dt)) / (2 * eps)
    return A, B


def exact_jacobian(w, mom, is_angle, x, u, dt):
    """Derivative of MBRL.ode by the chain rule in float64 (the network is piecewise linear): the quantity the
    central differences approximate and the tensor-core kernel computes in bf16."""
    x, u = np.asarray(x, dtype=np.float64), np.asarray(u, dtype=np.float64)
    R, n = x.shape
    m, nh = u.shape[1], n // 2
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    o = (observation(x, is_angle) - f64(mom["o_mean"])) / f64(mom["o_var"])
    uu = (u - f64(mom["u_mean"])) / f64(mom["u_var"])
    h = np.concatenate([o, uu], axis=1)
    z1 = h @ f64(w["W1"]).T + f64(w["b1"])
    z2 = np.maximum(z1, 0) @ f64(w["W2"]).T + f64(w["b2"])
    A, B = np.zeros((R, n, n)), np.zeros((R, n, m))
    for i in range(nh):
        A[:, i, nh + i] = 1.0
    n_obs = o.shape[1]
    for j in range(nh):
        g2 = f64(w["W3"])[j][None, :] * (z2 > 0)
        g1 = (g2 @ f64(w["W2"])) * (z1 > 0)
        gin = g1 @ f64(w["W1"]) * f64(mom["dx_var"])[j] / dt        # d rhs[nh + j] / d h
        c = 0
        for i, ang in enumerate(is_angle):
            if ang:
                A[:, nh + j, i] = gin[:, c] * (-np.sin(x[:, i])) / f64(mom["o_var"])[c] + gin[:, c + 1] * np.cos(x[:, i]) / f64(mom["o_var"])[c + 1]
                c += 2
            else:
                A[:, nh + j, i] = gin[:, c] / f64(mom["o_var"])[c]
                c += 1
        for jj in range(m):
            B[:, nh + j, jj] = gin[:, n_obs + jj] / f64(mom["u_var"])[jj]
    return A, B


class QuadCost:
    """c(x, u) = sum q_i x_i^2 + r u^2 (running), cN(x) = sum qN_i x_i^2 (final): the cost family of the examples
    (examples/mbrl/cart_pole_hpc.py:24-33), with its expansion as iLQR.cost_init derives it (ilqr.py:535-583)."""

    def __init__(self, q, r, qN):
        self.q, self.r, self.qN = np.asarray(q, float), np.asarray(r, float), np.asarray(qN, float)

    def c(self, x, u):
        return np.sum(self.q * x ** 2, axis=-1) + np.sum(self.r * u ** 2, axis=-1)

    def cN(self, x):
        return np.sum(self.qN * x ** 2, axis=-1)


def ilqr_backward(cost, A, B, xx, uu, mu, dt, reg_type=2):
    """iLQR.backward_pass (ilqr.py:233-346), unconstrained, one trajectory: A [T,n,n], B [T,n,m] continuous-time."""
    T, n = uu.shape[0], xx.shape[1]
    m = uu.shape[1]
    x = xx[-1]
    vx = (2 * cost.qN * x).reshape(n, 1) * dt
    Vxx = np.diag(2 * cost.qN) * dt
    KK, kk, V1, V2 = [], [], [0.0], [0.0]
    for i in range(T - 1, -1, -1):
        x, u = xx[i], uu[i]
        Cxx, Cuu, Cxu = np.diag(2 * cost.q) * dt, np.diag(2 * cost.r) * dt, np.zeros((n, m))
        cx, cu = (2 * cost.q * x).reshape(n, 1) * dt, (2 * cost.r * u).reshape(m, 1) * dt
        Ad, Bd = A[i] * dt + np.eye(n), B[i] * dt
        qx = cx + Ad.T @ vx
        qu = cu + Bd.T @ vx
        Qxx = Cxx + Ad.T @ (Vxx @ Ad)
        Quu = Cuu + Bd.T @ (Vxx @ Bd)
        Qux = Cxu.T + Bd.T @ (Vxx @ Ad)
        VxxReg = Vxx + mu * np.eye(n) * (reg_type == 1)
        QuuReg = Cuu + Bd.T @ (VxxReg @ Bd) + mu * np.eye(m) * (reg_type == 2)
        QuxReg = Cxu.T + Bd.T @ (VxxReg @ Ad)
        try:
            np.linalg.cholesky(QuuReg)
        except np.linalg.LinAlgError:
            return V1, V2, False, KK, kk
        kt = -np.linalg.solve(QuuReg, qu)
        Kt = -np.linalg.solve(QuuReg, QuxReg)
        vx = qx + Kt.T @ Quu @ kt + Kt.T @ qu + Qux.T @ kt
        Vxx = Qxx + Kt.T @ Quu @ Kt + Kt.T @ Qux + Qux.T @ Kt
        Vxx = 0.5 * (Vxx + Vxx.T)
        V2.insert(0, (0.5 * kt.T @ Quu @ kt).item())
        V1.insert(0, (kt.T @ qu).item())
        KK.insert(0, Kt)
        kk.insert(0, kt)
    return V1, V2, True, KK, kk


def traj_cost(cost, xx, uu, dt):
    """sum_k c(x_k, u_k) dt + cN(x_T) dt  (fast_step's transition costs, ilqr.py:213-224)."""
    return float(np.sum(cost.c(xx[:-1], uu)) * dt + cost.cN(xx[-1]) * dt)


def ilqr_solve(w, mom, is_angle, cost, x0, T, dt, max_iters=20, tol_fun=1e-7, tol_grad=1e-4, linearize="fd", trace=None):
    """iLQR.run_optim (ilqr.py:399-485) for ONE trajectory through the learned dynamics (fastForward=True,
    constrained=False, serial line search).  linearize: "fd" (the reference) or "exact"."""
    n, m = len(x0), mom["u_mean"].shape[0]
    lin = fd_linearize if linearize == "fd" else exact_jacobian
    X, U = rollout(w, mom, is_angle, np.asarray(x0, float)[None], dt, T=T)
    xx, uu = X[0], U[0]
    J = traj_cost(cost, xx, uu, dt)
    mu, mu_d, mu_d0, mu_min, mu_max = 1e-6, 1.0, 1.6, 1e-6, 1e10
    alphas = 10 ** np.linspace(0, -3, 11)
    success_gradient = False
    res = {"cost0": J, "xx0": xx.copy(), "costs": [], "mus": []}
    KK = kk = None
    it = 0

    def inc():
        nonlocal mu, mu_d
        mu_d = max(mu_d0, mu_d0 * mu_d)
        mu = max(mu_min, mu * mu_d)

    def dec():
        nonlocal mu, mu_d
        mu_d = min(mu_d / mu_d0, 1 / mu_d0)
        mu = mu * mu_d * (mu > mu_min)

    for it in range(1, max_iters + 1):
        A, B = lin(w, mom, is_angle, xx[:-1], uu, dt)
        V1, V2, ok, KKc, kkc = ilqr_backward(cost, A, B, xx, uu, mu, dt)
        if it == 1:
            res.update(A0=A, B0=B, KK1=np.array(KKc), kk1=np.array(kkc).reshape(len(kkc), m), sV1=np.sum(V1), sV2=np.sum(V2))
        success_fw = False
        if not ok:
            inc()
        else:
            g_norm = np.mean(np.max(np.abs(np.array(kkc).reshape(len(kkc), m) / (np.abs(uu) + 1)), axis=0))
            if g_norm < tol_grad and mu < 1e-5:
                dec()
                success_gradient = True
            cands = []
            for alpha in alphas:
                Xc, Uc = rollout(w, mom, is_angle, np.asarray(x0, float)[None], dt,
                                 gains=(np.array(KKc)[None], np.array(kkc).reshape(1, len(kkc), m), xx[None], uu[None]), alpha=alpha)
                c = traj_cost(cost, Xc[0], Uc[0], dt)
                cands.append((c, Xc[0], Uc[0]))
                if it == 1 and alpha == 1.0:
                    res["cost_alpha1"] = c
                if np.any(Xc[0] > 1e8):
                    break
                dcost = J - c
                expected = -alpha * (np.sum(V1) + alpha * np.sum(V2))
                z = dcost / expected if expected > 0 else np.sign(dcost)
                if z > 0:
                    success_fw = True
                    break
        if success_fw:
            best = int(np.argmin([c[0] for c in cands]))
            J, xx, uu = cands[best]
            KK, kk = KKc, kkc
            dec()
            res["costs"].append(J)
            res["mus"].append(mu)
            if dcost < tol_fun or success_gradient:
                break
        else:
            inc()
            res["costs"].append(J)
            res["mus"].append(mu)
            if mu > mu_max:
                break
    res.update(cost=J, xx=xx, uu=uu, iters=it)
    return res
