"""The oracle (oracle/pygent_oracle.c) against outputs of the UNMODIFIED reference (tests/golden/*.npz,
made by oracle/make_golden.py).  This is what pins the oracle; the GPU tests then compare the CUDA
path with the oracle.  CPU only."""
import numpy as np
import pytest

import oracle
from cases import CASES, ILQR_RUNS, oracle_problem, rk45_tol


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1.0))) if a.size else 0.0


SYSTEMS = sorted({c[2] for c in CASES.values()}) + ["car", "marbot", "angular_pendulum"]


@pytest.mark.parametrize("system", SYSTEMS)
def test_dynamics_and_jacobians_match_reference(golden, system):
    g = golden("models")
    xs, us = g[system + "/x"], g[system + "/u"]
    f = np.array([oracle.f(system, x, u) for x, u in zip(xs, us)])
    assert rel(f, g[system + "/f"]) < 2e-12
    if system + "/A" in g.files:
        AB = [oracle.AB(system, x, u) for x, u in zip(xs, us)]
        assert rel(np.array([a for a, _ in AB]), g[system + "/A"]) < 2e-11
        assert rel(np.array([b for _, b in AB]), g[system + "/B"]) < 2e-11


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("tag,S,integ", [("rk4x1", 1, 0), ("rk4x4", 4, 0), ("euler", 1, 1)])
def test_fixed_step_rollout_matches_reference(golden, case, tag, S, integ):
    g = golden("steps")
    p = oracle_problem(case, S=S, integrator=integ)
    r = oracle.rollout(p, g[case + "/x0"][None], U=g[case + "/U"][None])
    if "%s/%s/X" % (case, tag) not in g.files:
        pytest.skip("the reference's own fast_step fails on MarBot's list-valued ode (environments.py:314): no Euler golden")
    Xref, cref = g["%s/%s/X" % (case, tag)], g["%s/%s/c" % (case, tag)]
    assert rel(r["X"][0], Xref) < 1e-10
    assert abs(r["cost"][0] - cref.sum()) <= 1e-10 * (abs(cref.sum()) + 1)
    assert bool(r["terminated"][0]) == bool(g["%s/%s/terminated" % (case, tag)])


@pytest.mark.parametrize("case", sorted(CASES))
def test_rk4x4_tracks_native_rk45(golden, case):
    """north_star tolerance: 1e-6 relative against the reference's own adaptive integrator.  The
    torque/force-input double pendulums are stiffer: there the gap is the error of the reference's
    RK45 at its default rtol=1e-3, hence the looser bound."""
    g = golden("steps")
    p = oracle_problem(case, S=4)
    r = oracle.rollout(p, g[case + "/x0"][None], U=g[case + "/U"][None])
    tol = rk45_tol(case)
    assert rel(r["X"][0], g[case + "/rk45/X"]) < tol


def box(case, kw):
    """oracle.ilqr_params arguments of a `constrained=True` run: the control limits are the environment's uMax (ilqr.py:310)."""
    if not kw.get("constrained"):
        return {}
    from pygent_b200.examples import make_env
    return {"constrained": True, "lims": [float(v) for v in np.ravel(make_env(case).uMax)]}


@pytest.mark.parametrize("run", sorted(ILQR_RUNS))
def test_ilqr_first_iteration_matches_reference(golden, run):
    case, T, S, integ, kw = ILQR_RUNS[run]
    g = golden(run)
    p = oracle_problem(case, S=S, integrator=integ)
    q = oracle.ilqr_params(parallel=kw.get("parallel", False), reg_type=kw.get("regType", 2),
                           finite_diff=kw.get("finite_diff", False), **box(case, kw))
    r0 = oracle.rollout(p, g["x0"][None], T=T, add_final_cost=True)
    assert rel(r0["X"][0], g["xx0"]) < 1e-10
    assert abs(r0["cost"][0] - g["cost0"]) < 1e-10 * abs(g["cost0"])
    KK, kk, s1, s2, gn, ok = oracle.ilqr_backward(p, q, g["xx0"], g["uu0"], 1e-6)
    assert ok
    # finite_diff: second differences with eps = 1e-3 amplify the last-bit differences between two evaluations of the
    # same cost expression (numpy there, C/CUDA here) by 1/(4 eps^2) -- 1e-10 absolute on the Hessians, 1e-7 on gains
    tol = 1e-6 if kw.get("finite_diff") else 1e-8
    assert rel(KK, g["bw_KK"]) < tol
    assert rel(kk, g["bw_kk"]) < tol
    assert rel([s1, s2], g["bw_sumV"]) < tol / 10
    xo, uo, c, div, term = oracle.ilqr_forward(p, q, 1.0, g["x0"], g["xx0"], g["uu0"], g["bw_KK"], g["bw_kk"])
    assert rel(xo, g["fw1_xx"]) < 1e-9
    assert rel(uo, g["fw1_uu"]) < 1e-9
    assert abs(c - g["fw1_cost"]) < 1e-9 * abs(g["fw1_cost"])


@pytest.mark.parametrize("run", sorted(ILQR_RUNS))
def test_ilqr_solve_matches_reference(golden, run):
    """Whole run_optim: same iteration count, same per-iteration cost / mu trace, same result."""
    case, T, S, integ, kw = ILQR_RUNS[run]
    g = golden(run)
    p = oracle_problem(case, S=S, integrator=integ)
    q = oracle.ilqr_params(parallel=kw.get("parallel", False), reg_type=kw.get("regType", 2),
                           finite_diff=kw.get("finite_diff", False), **box(case, kw))
    out = oracle.ilqr_solve(p, q, g["x0"][None], T, trace=True)
    iters = int(g["iters"])
    plateau = out["iters"][0] != iters
    if plateau:
        # allowed only on a converged plateau: the reference spent its last iterations at the optimum (cost unchanged to
        # 1e-13) raising mu after rejected line searches, where accept / reject is the sign of a difference at rounding level
        # (dcost ~ 1e-13 against an expected reduction of the same size) -- ilqr_cart_pole_constrained: 45 there, 46 here
        gc = g["trace"][:, 0]
        assert abs(int(out["iters"][0]) - iters) <= 1 and np.ptp(gc[-8:]) < 1e-12 * gc[-1]
        assert abs(out["cost"][0] - g["cost"]) < 1e-12 * abs(g["cost"])
        iters = min(iters, int(out["iters"][0]))
    tr = out["trace"][0]
    # golden trace rows: (cost, mu, alpha) at the START of each iteration; ours: after each iteration
    ours_cost = np.concatenate([[out["cost0"][0]], tr[: iters - 1, 0]])
    w = 100 if kw.get("finite_diff") else 1  # finite-difference noise of the cost Hessians, see the first-iteration test
    assert rel(ours_cost, g["trace"][:iters, 0]) < 1e-7 * w
    ours_mu = np.concatenate([[1e-6], tr[: iters - 1, 2]])
    assert rel(ours_mu, g["trace"][:iters, 1]) < 1e-12
    assert abs(out["cost"][0] - g["cost"]) < 1e-7 * w * abs(g["cost"])
    assert rel(out["xx"][0], g["xx"]) < 1e-5 * w
    assert rel(out["uu"][0], g["uu"]) < 1e-5 * w
    if float(g["mu"]) <= 1e10 and not plateau:
        assert rel(out["KK"][0], g["KK"]) < 1e-4 * w
        assert out["alpha"][0] == pytest.approx(float(g["alpha"]))
    # else: the reference left through "mu > mu_max" on its last iteration -- a rounding-level
    # knife edge at mu ~ 1e9 (dcost ~ 1e-12); costs and trajectories above still agree


# ---- learned-MLP dynamics: the numpy restatement (oracle/mlp_oracle.py) against the reference's own classes ---------
MLP_IS_ANGLE = [False, True, False, False]


def _mlp_setup():
    from oracle import mlp_oracle as mo
    return mo, mo.make_weights(6, 2), mo.make_moments(5, 1, 2), mo.QuadCost([10, 5, .01, .01], [0.1], [100, 100, 10, 10])


def test_mlp_oracle_matches_reference_ode_and_euler(golden):
    mo, w, mom, _ = _mlp_setup()
    g = golden("mlp_cart_pole")
    assert rel(mo.mbrl_ode(w, mom, MLP_IS_ANGLE, g["xs"], g["us"], float(g["dt"])), g["rhs"]) < 2e-6
    X, _ = mo.rollout(w, mom, MLP_IS_ANGLE, g["x0"][None], float(g["dt"]), U=g["U"][None])
    assert rel(X[0], g["X"]) < 2e-6


def test_mlp_ilqr_oracle_matches_reference(golden):
    """iLQR(StateSpaceModel(MBRL.ode), fastForward=True) of the reference (oracle/make_golden_mlp.py).  The reference
    linearises by central differences through its fp32 network: eps = 1e-3 against ~1e-7 rounding, so Jacobians carry
    ~1e-4 relative noise that depends on the summation order of the GEMM -- gains agree to that noise, costs much
    better."""
    mo, w, mom, cost = _mlp_setup()
    g = golden("mlp_ilqr_cart_pole")
    T, dt = int(g["T"]), float(g["dt"])
    r = mo.ilqr_solve(w, mom, MLP_IS_ANGLE, cost, g["x0"], T, dt, max_iters=12)
    assert rel(r["cost0"], g["cost0"]) < 1e-7 and rel(r["xx0"], g["xx0"]) < 1e-6
    assert rel(r["A0"] * dt + np.eye(4), g["Ad0"]) < 5e-5 and rel(r["B0"] * dt, g["Bd0"]) < 5e-5
    assert rel(r["KK1"], g["KK1"]) < 2e-2 and rel(r["kk1"], g["kk1"]) < 2e-2
    assert rel(r["sV1"], g["sV1"]) < 1e-3 and rel(r["sV2"], g["sV2"]) < 1e-3
    assert rel(r["cost_alpha1"], g["cost_alpha1"]) < 1e-5
    assert rel(r["cost"], g["cost"]) < 1e-4
    # the exact derivative of the piecewise-linear network is what the differences approximate
    Ae, Be = mo.exact_jacobian(w, mom, MLP_IS_ANGLE, g["xx0"][:-1], np.zeros((T, 1)), dt)
    assert rel(Ae, r["A0"]) < 5e-2 and rel(Be, r["B0"]) < 5e-3
