"""CUDA path vs the oracle and vs the reference's golden outputs.  Everything here goes through the
C ABI of the compiled problem libraries (pygent_b200/lib.py -> include/pygent_b200.h)."""
import numpy as np
import pytest
import torch

import oracle
from cases import CASES, ILQR_RUNS, oracle_problem, rk45_tol
from pygent_b200 import lib as pglib
from pygent_b200.examples import make_env, make_ilqr

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1.0))) if a.size else 0.0


def dev(a, dtype=torch.float64):
    """host [B, ...] -> device [..., B]"""
    return torch.as_tensor(np.ascontiguousarray(np.moveaxis(np.asarray(a, dtype=float), 0, -1)), dtype=dtype, device="cuda")


def dev_U(U, dtype=torch.float64):
    """host [B, T, m] -> device [T, m, B]"""
    return torch.as_tensor(np.ascontiguousarray(np.asarray(U, dtype=float).transpose(1, 2, 0)), dtype=dtype, device="cuda")


def host(t):
    """device [..., B] -> host [B, ...]"""
    return np.moveaxis(t.detach().to("cpu", torch.float64).numpy(), -1, 0)


def solver(case, T, x0=None, **kw):
    return make_ilqr(case, T, x0=x0, **kw)


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", sorted(CASES))
def test_generated_functions_match_oracle(case):
    """f, A, B and the cost expansion of the generated device code at random points."""
    alg = solver(case, 10)
    L = alg.lib
    n, m = L.n, L.m
    rng = np.random.default_rng(1)
    P = 257
    xs, us, ts = rng.uniform(-2.5, 2.5, (P, n)), rng.uniform(-3, 3, (P, m)), rng.uniform(0, 2, P)
    if case == "triple":
        xs[:, 4:] *= 0.3
    r = L.eval(dev(xs), dev(us), torch.as_tensor(ts, device="cuda"))
    system = CASES[case][2]
    f = np.array([oracle.f(system, x, u) for x, u in zip(xs, us)])
    assert rel(host(r["f"]), f) < 1e-11
    AB = [oracle.AB(system, x, u, fd=not L.has_ab) for x, u in zip(xs, us)]
    assert rel(host(r["A"]), np.array([a for a, _ in AB])) < 1e-9
    assert rel(host(r["B"]), np.array([b for _, b in AB])) < 1e-9
    cp = CASES[case][3]()[2]
    qx, qf = np.array(cp["qx"]), np.array(cp["qf"])
    lx, lf = np.array(cp.get("lx", [0.0] * n)), np.array(cp.get("lf", [0.0] * n))   # affine part (angular pendulum)
    ru = np.array(cp["ru"])[None] + np.array(cp.get("re", [0.0] * m))[None] * np.exp(-cp.get("rd", 0.0) * ts)[:, None]
    assert rel(host(r["c"]), cp.get("c0", 0.0) + (qx * xs ** 2 + lx * xs).sum(1) + (ru * us ** 2).sum(1)) < 1e-12
    assert rel(host(r["cx"]), 2 * qx * xs + lx) < 1e-12
    assert rel(host(r["cu"]), 2 * ru * us) < 1e-12
    assert rel(host(r["Cxx"]), np.broadcast_to(np.diag(2 * qx), (P, n, n))) < 1e-12
    assert rel(host(r["Cuu"])[:, 0, 0], 2 * ru[:, 0]) < 1e-12
    assert rel(host(r["Cxu"]), 0) == 0
    assert rel(host(r["cf"]), cp.get("cf0", 0.0) + (qf * xs ** 2 + lf * xs).sum(1)) < 1e-12
    assert rel(host(r["cfx"]), 2 * qf * xs + lf) < 1e-12
    assert rel(host(r["Cfxx"]), np.broadcast_to(np.diag(2 * qf), (P, n, n))) < 1e-12


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("S,integ", [(1, 0), (4, 0), (1, 1)])
def test_rollout_matches_oracle(case, S, integ):
    alg = solver(case, 10)
    L = alg.lib
    n, m = L.n, L.m
    rng = np.random.default_rng(2)
    B, T = 96, 60
    x0 = np.array(CASES[case][4]) + rng.uniform(-0.2, 0.2, (B, n))
    umax = float(np.ravel(alg.environment.uMax)[0])
    U = rng.uniform(-umax, umax, (B, T, m)) * 0.5
    dt = CASES[case][5]
    r = L.rollout(dev(x0), dev_U(U), S=S, dt=dt, integrator=integ, history=True, add_final_cost=True)
    p = oracle_problem(case, S=S, integrator=integ)
    ro = oracle.rollout(p, x0, U=U, add_final_cost=True)
    X = np.moveaxis(r["X"].cpu().numpy(), -1, 0)  # [B, T+1, n]
    assert rel(X, ro["X"]) < 1e-9
    assert rel(host(r["xN"]), ro["xN"]) < 1e-9
    assert rel(r["cost"].cpu().numpy(), ro["cost"]) < 1e-9
    assert np.array_equal(r["terminated"].cpu().numpy(), ro["terminated"])


@pytest.mark.parametrize("case", sorted(CASES))
def test_rollout_matches_reference_golden(golden, case):
    """Same inputs as the unmodified reference: RK4x4 (1e-10), RK4x1, Euler, and native RK45 (1e-6)."""
    g = golden("steps")
    alg = solver(case, 10)
    L = alg.lib
    dt = CASES[case][5]
    x0 = dev(g[case + "/x0"][None])
    U = torch.as_tensor(np.ascontiguousarray(g[case + "/U"][:, :, None]), device="cuda")
    for tag, S, integ, tol in (("rk4x4", 4, 0, 1e-10), ("rk4x1", 1, 0, 1e-10), ("euler", 1, 1, 1e-10), ("rk45", 4, 0, None)):
        if "%s/%s/X" % (case, tag) not in g.files:
            continue   # MarBot: the reference's own fast_step fails on its list-valued ode, no Euler golden
        r = L.rollout(x0, U, S=S, dt=dt, integrator=integ, history=True)
        X = r["X"].cpu().numpy()[:, :, 0]
        if tol is None:
            tol = rk45_tol(case)
        assert rel(X, g["%s/%s/X" % (case, tag)]) < tol, tag
        if tag != "rk45":
            c = g["%s/%s/c" % (case, tag)].sum()
            assert abs(r["cost"].item() - c) < 1e-10 * (abs(c) + 1)


def test_rollout_edge_cases():
    alg = solver("cart_pole", 10)
    L = alg.lib
    # empty batch and zero-length horizon
    r = L.rollout(torch.zeros((4, 0), dtype=torch.float64, device="cuda"), T=5, S=4, dt=0.02)
    assert r["xN"].shape == (4, 0)
    x0 = dev(np.array([[0.1, 3.0, 0.0, 0.2]] * 3))
    r = L.rollout(x0, T=0, S=4, dt=0.02, history=True)
    assert torch.equal(r["xN"], x0) and torch.equal(r["X"][0], x0) and float(r["cost"].abs().max()) == 0.0
    # ragged batch (not a multiple of the CTA size) and u == 0 when U is NULL
    B = 131
    rng = np.random.default_rng(3)
    xs = np.array([0, np.pi, 0, 0]) + rng.uniform(-0.1, 0.1, (B, 4))
    r1 = L.rollout(dev(xs), T=20, S=4, dt=0.02)
    r2 = L.rollout(dev(xs), torch.zeros((20, 1, B), dtype=torch.float64, device="cuda"), S=4, dt=0.02)
    assert torch.equal(r1["xN"], r2["xN"]) and torch.equal(r1["cost"], r2["cost"])
    # termination box and terminal cost
    xs = np.array([[1.19, np.pi, 3.0, 0.0], [0.0, np.pi, 0.0, 0.0]])
    r = L.rollout(dev(xs), T=3, S=4, dt=0.02, terminal_cost=7.0)
    r0 = L.rollout(dev(xs), T=3, S=4, dt=0.02, terminal_cost=0.0)
    assert r["terminated"].cpu().tolist() == [1, 0]
    extra = (r["cost"] - r0["cost"]).cpu().numpy()
    assert extra[1] == 0 and extra[0] == pytest.approx(7.0 * 0.02 * 3, rel=1e-12)
    # wrong device / dtype / shape fail loudly
    with pytest.raises(pglib.PygentB200Error):
        L.rollout(torch.zeros((4, 2), dtype=torch.float64), T=1)
    with pytest.raises(pglib.PygentB200Error):
        L.rollout(torch.zeros((5, 2), dtype=torch.float64, device="cuda"), T=1)
    with pytest.raises(TypeError):
        L.rollout(torch.zeros((4, 2), dtype=torch.float16, device="cuda"), T=1)


@pytest.mark.parametrize("B,T,dtype", [(65536, 500, torch.float64), (19000, 130, torch.float64), (40001, 77, torch.float32),
                                       (40001, 77, torch.float64), (33, 60, torch.float64), (75776, 100, torch.float64)])
def test_balanced_rollout_is_bit_identical(B, T, dtype):
    """pg_rollout_balanced (persistent workers + ticket queue, time chunks of 25) must return exactly the bits of
    pg_rollout: states, costs, termination flags, history; also when continuing an accumulated cost."""
    alg = solver("cart_pole", 10)
    L = alg.lib
    gen = torch.Generator(device="cpu").manual_seed(B)
    x0 = torch.zeros((4, B), dtype=torch.float64)
    x0[0] = (torch.rand(B, generator=gen, dtype=torch.float64) - 0.5) * 2.0
    x0[1] = np.pi + (torch.rand(B, generator=gen, dtype=torch.float64) - 0.5)
    U = (torch.rand((T, 1, B), generator=gen, dtype=torch.float64) * 2 - 1) * 10.0
    x0d, Ud = x0.to("cuda", dtype), U.to("cuda", dtype)
    hist = B <= 20000
    a = L.rollout(x0d, Ud, S=4, dt=0.02, history=hist, add_final_cost=True, terminal_cost=3.0, balanced=False)
    b = L.rollout(x0d, Ud, S=4, dt=0.02, history=hist, add_final_cost=True, terminal_cost=3.0, balanced=True)
    if dtype == torch.float32:  # two different kernels: fp32 FMA contraction may differ, compare to rounding
        same = lambda p, q: bool(torch.allclose(p, q, rtol=2e-3, atol=2e-3))
    else:
        same = torch.equal
    assert same(a["xN"], b["xN"]) and same(a["cost"], b["cost"]) and torch.equal(a["terminated"], b["terminated"])
    if hist:
        assert same(a["X"], b["X"])
    assert all(int(w.item()) == 0 for w in L.queue_status_words())  # no worker ever gave up waiting
    L.check_queues()
    # continuation of an accumulated cost (the host-pipelined path) and the Euler integrator
    c0 = torch.full((B,), 1.5, dtype=dtype, device="cuda")
    a2 = L.rollout(a["xN"], Ud, S=4, dt=0.02, t0=1.0, integrator=1, accumulate_cost=True, out={"cost": c0.clone()}, balanced=False)
    b2 = L.rollout(a["xN"], Ud, S=4, dt=0.02, t0=1.0, integrator=1, accumulate_cost=True, out={"cost": c0.clone()}, balanced=True)
    assert same(a2["xN"], b2["xN"]) and same(a2["cost"], b2["cost"])


@pytest.mark.parametrize("case,B,T", [("cart_pole", 40001, 130), ("car", 777, 33), ("cart_pole", 64, 7)])
def test_in_kernel_action_stream(case, B, T):
    """"in-kernel Philox with the same counter scheme, validated against the host stream".
    (1) pg_random_actions on the device == pygent_b200/random_actions.py on the host, bit for bit (m = 1 and m = 2, with
    offsets); (2) a rollout that draws its controls in the kernel (pg_rollout_random, plain and ticket-queue kernel) ==
    the rollout over the materialised U, bit for bit, fp64 and fp32."""
    from pygent_b200.random_actions import random_actions
    alg = solver(case, 10)
    L = alg.lib
    m, n = L.m, L.n
    u_max = [10.0, 0.7][:m]
    seed = 0x1234_5678_9ABC_DEF0 + B
    U = L.random_actions(B, T, seed, u_max)
    Uh = random_actions(B, T, seed, u_max)
    assert np.array_equal(U.cpu().numpy(), Uh)
    assert np.array_equal(L.random_actions(B, 5, seed, u_max, k_offset=T - 5).cpu().numpy(), Uh[T - 5:])
    rng = np.random.default_rng(B)
    x0 = torch.as_tensor(np.array(CASES[case][4])[:, None] + rng.uniform(-0.2, 0.2, (n, B)), device="cuda")
    for dtype in (torch.float64, torch.float32):
        x0d, Ud = x0.to(dtype), L.random_actions(B, T, seed, u_max, dtype=dtype)
        if dtype == torch.float32:
            assert np.array_equal(Ud.cpu().numpy(), random_actions(B, T, seed, u_max, dtype=np.float32))
        for balanced in (False, True):
            a = L.rollout(x0d, Ud, S=4, dt=0.02, history=B < 1000, balanced=balanced)
            b = L.rollout(x0d, None, T=T, S=4, dt=0.02, history=B < 1000, balanced=balanced, actions=(seed, u_max))
            assert torch.equal(a["xN"], b["xN"]) and torch.equal(a["cost"], b["cost"]) and torch.equal(a["terminated"], b["terminated"])
            if B < 1000:
                assert torch.equal(a["X"], b["X"])
    # a horizon split into two launches continues the same stream (k_offset)
    a = L.rollout(x0, None, T=T, S=4, dt=0.02, balanced=False, actions=(seed, u_max))
    h = T // 2
    p1 = L.rollout(x0, None, T=h, S=4, dt=0.02, balanced=False, actions=(seed, u_max))
    p2 = L.rollout(p1["xN"], None, T=T - h, S=4, dt=0.02, t0=h * 0.02, balanced=False, accumulate_cost=True, out={"cost": p1["cost"].clone()},
                   actions=(seed, u_max, h))
    assert torch.equal(a["xN"], p2["xN"])
    L.check_queues()


def test_queue_worker_timeout_is_reported(monkeypatch):
    """a worker of the ticket-queue kernels that gives up waiting (~2 s) must not go unnoticed.  PG_QUEUE_FAULT=1
    makes the launch skip the first-chunk tickets, so every worker waits for a unit nobody publishes; the kernel still ends
    (no hang, no trap), the sticky status word of the workspace is set, and the next host-side check raises."""
    alg = solver("cart_pole", 10)
    L = alg.lib
    B, T = 4096, 60
    x0 = torch.zeros((4, B), dtype=torch.float64, device="cuda")
    U = torch.zeros((T, 1, B), dtype=torch.float64, device="cuda")
    L.rollout(x0, U, S=4, dt=0.02, balanced=True)
    L.check_queues()
    monkeypatch.setenv("PG_QUEUE_FAULT", "1")
    L.rollout(x0, U, S=4, dt=0.02, balanced=True)
    monkeypatch.delenv("PG_QUEUE_FAULT")
    L.rollout(x0, U, S=4, dt=0.02, balanced=True)  # a later good launch on the same workspace does not clear the word
    with pytest.raises(pglib.PygentB200Error, match="timed out"):
        L.check_queues()
    L.check_queues()  # reported once, then cleared
    r = L.rollout(x0, U, S=4, dt=0.02, balanced=True)
    assert torch.equal(r["xN"], L.rollout(x0, U, S=4, dt=0.02, balanced=False)["xN"])


@pytest.mark.parametrize("run", sorted(ILQR_RUNS))
def test_ilqr_backward_and_forward_match_golden(golden, run):
    """First backward pass (K, k, sum V1, sum V2) and the alpha = 1 forward pass of the reference."""
    case, T, S, integ, kw = ILQR_RUNS[run]
    g = golden(run)
    alg = solver(case, T, x0=g["x0"], env_kw={"substeps": S}, fastForward=bool(integ), regType=kw.get("regType", 2),
                 finite_diff=kw.get("finite_diff", False), constrained=kw.get("constrained", False))
    assert rel(alg.xx, g["xx0"]) < 1e-10
    assert alg.cost == pytest.approx(float(g["cost0"]), rel=1e-11)
    V1, V2, ok, KK, kk = alg.backward_pass()
    assert ok
    # finite_diff: second differences with eps = 1e-3 amplify the last-bit differences between two evaluations of the
    # same cost expression (numpy there, CUDA with FMA here) by 1/(4 eps^2) -- 1e-10 absolute on the Hessians
    tol = 1e-6 if kw.get("finite_diff") else 1e-8
    assert rel(KK, g["bw_KK"]) < tol
    assert rel(kk[..., 0], g["bw_kk"]) < tol
    assert rel([V1, V2], g["bw_sumV"]) < tol / 10
    xx, uu, cost, term = alg.forward_pass(1.0, g["bw_KK"], g["bw_kk"])
    assert rel(xx, g["fw1_xx"]) < 1e-9
    assert rel(uu, g["fw1_uu"]) < 1e-9
    assert cost == pytest.approx(float(g["fw1_cost"]), rel=1e-9)


@pytest.mark.parametrize("run", sorted(ILQR_RUNS))
def test_ilqr_solve_matches_reference(golden, run):
    """Whole run_optim of the unmodified reference: iteration count, final cost, trajectory, gains."""
    case, T, S, integ, kw = ILQR_RUNS[run]
    g = golden(run)
    alg = solver(case, T, x0=g["x0"], env_kw={"substeps": S}, fastForward=bool(integ), regType=kw.get("regType", 2),
                 parallel=kw.get("parallel", False), finite_diff=kw.get("finite_diff", False), constrained=kw.get("constrained", False))
    xx, uu, cost = alg.run_optim()
    plateau = alg.iters != int(g["iters"])
    if plateau:
        # allowed only on a converged plateau (see tests/test_oracle_golden.py): the reference's last iterations sit at the
        # optimum, cost unchanged to 1e-13, and accept / reject is the sign of a rounding-level difference
        gc = g["trace"][:, 0]
        assert abs(alg.iters - int(g["iters"])) <= 1 and np.ptp(gc[-8:]) < 1e-12 * gc[-1]
        assert cost == pytest.approx(float(g["cost"]), rel=1e-12)
    w = 100 if kw.get("finite_diff") else 1  # finite-difference noise of the cost Hessians, see the first-iteration test
    assert cost == pytest.approx(float(g["cost"]), rel=1e-7 * w)
    assert rel(xx, g["xx"]) < 1e-5 * w
    assert rel(uu, g["uu"]) < 1e-5 * w
    if float(g["mu"]) <= 1e10 and not plateau:
        assert rel(alg.KK, g["KK"]) < 1e-4 * w
        assert alg.current_alpha == pytest.approx(float(g["alpha"]))
        assert alg.mu == pytest.approx(float(g["mu"]), rel=1e-9)


def reference_select(alphas, parallel, tol_fun, tol_grad, st, bw_ok, gnorm, sV1, sV2, costs, div, max_iters):
    """The line-search / mu / stopping logic of iLQR.run_optim (ilqr.py:407-485) for ONE trajectory, in plain
    python, applied to the numbers the GPU produced (so there is no rounding ambiguity in the comparison)."""
    mu, mu_d, cost, dcost, sg, iters = st["mu"], st["mu_d"], st["cost"], st["dcost"], st["sg"], st["iters"] + 1

    def inc(mu, mu_d):
        mu_d = max(1.6, 1.6 * mu_d)
        return max(1e-6, mu * mu_d), mu_d

    def dec(mu, mu_d):
        mu_d = min(mu_d / 1.6, 1 / 1.6)
        return mu * mu_d * (mu > 1e-6), mu_d
    status, accept, alpha = 0, 0, st["alpha"]
    if not bw_ok:
        mu, mu_d = inc(mu, mu_d)
        mu, mu_d = inc(mu, mu_d)
        if mu > 1e10:
            status = 3
    else:
        if gnorm < tol_grad and mu < 1e-5:
            mu, mu_d = dec(mu, mu_d)
            sg = True
        ok, tried = False, 0
        for i, a in enumerate(alphas):
            tried = i + 1
            if div[i]:
                break
            dcost = cost - costs[i]
            expected = -a * (sV1 + a * sV2)
            z = dcost / expected if expected > 0 else np.sign(dcost)
            if z > 0:
                ok = True
                if not parallel:
                    break
        if ok:
            best = int(np.argmin(costs[:tried]))
            cost, alpha, accept = costs[best], alphas[best], 1
            mu, mu_d = dec(mu, mu_d)
            if dcost < tol_fun:
                status = 1
            elif sg:
                status = 2
        else:
            mu, mu_d = inc(mu, mu_d)
            if mu > 1e10:
                status = 3
    if status == 0 and iters >= max_iters:
        status = 4
    return dict(mu=mu, mu_d=mu_d, cost=cost, dcost=dcost, sg=sg, iters=iters, alpha=alpha), status, accept


@pytest.mark.parametrize("case,T,parallel,constrained,reg", [
    ("cart_pole", 100, False, False, 2), ("cart_pole", 100, True, False, 2), ("cart_pole", 60, False, True, 2),
    ("cart_pole", 60, False, False, 1), ("acrobot", 80, True, False, 2), ("double_parallel", 100, True, False, 2),
    ("double_serial", 60, False, False, 2), ("pendulum", 60, False, True, 2), ("cart_pole_tv", 100, False, False, 2),
    ("car", 60, False, False, 2), ("car", 60, True, True, 2), ("marbot", 60, False, True, 2),
    ("angular_pendulum", 60, False, True, 2)])   # m = 2 (2-D box QP), FD Jacobians, affine cost
def test_ilqr_iterations_in_lockstep_with_oracle(case, T, parallel, constrained, reg):
    """25 iterations of a batch, checked one iteration at a time: from the GPU's own nominal trajectory the
    oracle recomputes the backward pass (K, k, sum V1/V2, gradient norm) and all 11 line-search costs, and
    the reference's decision logic is replayed on the GPU's numbers (accepted alpha, mu, status -- exact)."""
    rng = np.random.default_rng(5)
    n = len(CASES[case][4])
    B = 24
    x0 = np.array(CASES[case][4]) + rng.uniform(-0.3, 0.3, (B, n))
    alg = solver(case, T, x0=x0, parallel=parallel, constrained=constrained, regType=reg, maxIters=25)
    p = oracle_problem(case, S=4)
    q = oracle.ilqr_params(parallel=parallel, constrained=constrained, lims=list(alg.lims), reg_type=reg)
    alg._sel = alg._select_params()
    alg._rearm()
    st = [dict(mu=1e-6, mu_d=1.0, cost=float(alg.cost[b]), dcost=0.0, sg=False, iters=0, alpha=1.0) for b in range(B)]
    status = np.zeros(B, dtype=int)
    for it in range(25):
        xx, uu, mu = alg.xx, alg.uu, alg.mu
        alg.iterate()
        bw = {k: v.cpu().numpy() for k, v in alg._bw.items()}
        fw = {k: v.cpu().numpy() for k, v in alg._fw.items()}
        KKc, kkc = host(alg._KKc), host(alg._kkc)
        for b in range(B):
            if status[b] != 0:
                continue
            KK, kk, s1, s2, gn, ok = oracle.ilqr_backward(p, q, xx[b], uu[b], mu[b])
            assert ok == bool(bw["success"][b])
            if ok:
                assert rel(KKc[b], KK) < 1e-6 and rel(kkc[b], kk) < 1e-6
                assert rel([bw["dV"][0, b], bw["dV"][1, b]], [s1, s2]) < 1e-7
                assert rel(bw["gnorm"][b], gn) < 1e-7
                for i, a in enumerate(alg.alphas):
                    _, _, c, div, _ = oracle.ilqr_forward(p, q, a, x0[b], xx[b], uu[b], KKc[b], kkc[b])
                    assert div == bool(fw["diverged"][i, b])
                    if not div:
                        # candidates that blow up (cost orders of magnitude above the nominal) are unstable
                        # closed loops: rounding differences are amplified there, and they are never accepted
                        big = 50 * (abs(st[b]["cost"]) + 1)
                        if c < big:
                            assert abs(fw["cost"][i, b] - c) < 1e-8 * (abs(c) + 1), (it, b, i)
                        else:  # both must agree that this candidate is hopeless; its digits are chaos
                            assert fw["cost"][i, b] > 0.2 * big, (it, b, i)
            st[b], status[b], accept = reference_select(
                alg.alphas, parallel, alg.tolFun, alg.tolGrad, st[b], ok, bw["gnorm"][b], bw["dV"][0, b], bw["dV"][1, b],
                fw["cost"][:, b], fw["diverged"][:, b], 25)
        assert np.array_equal(alg.status, status), it
        assert np.array_equal(alg.iters, [s["iters"] for s in st])
        assert np.array_equal(alg.mu, [s["mu"] for s in st]), it
        assert np.array_equal(alg.cost, [s["cost"] for s in st])
        # the committed nominal trajectory is the winning candidate (recomputed): its cost must be reproducible
        if it % 8 == 0:
            r = alg.lib.rollout(alg._x0, alg._ubar, S=4, dt=alg.dt, add_final_cost=True, history=True)
            assert rel(r["cost"].cpu().numpy(), alg.cost) < 1e-11
            assert rel(r["X"].cpu().numpy(), alg._xbar.cpu().numpy()) < 1e-11  # other kernel: FMA contraction may differ


def free_running_agreement(alg, so, cost_tol, max_iter_diff):
    """A GPU solve against the oracle's solve of the same starts, both running free.  Every iteration is checked in lock
    step elsewhere (test_ilqr_iterations_in_lockstep_with_oracle: the GPU's numbers match the oracle's to 1e-8 and its
    decisions are the reference's decisions on those numbers); free-running, the two differ in rounding (the GPU contracts
    to FMA, the oracle is compiled with -ffp-contract=off), which can only move a decision whose margin is at rounding
    level: the `dcost < tolFun` stopping test near the optimum, i.e. WHEN a trajectory stops, not WHERE.  So: every final
    cost agrees to `cost_tol`, every iteration count to within `max_iter_diff`, and no trajectory ends above its start."""
    rc = np.abs(alg.cost - so["cost"]) / np.abs(so["cost"])
    assert np.all(rc <= cost_tol), (rc.max(), np.where(rc > cost_tol)[0])
    di = np.abs(alg.iters - so["iters"])
    assert di.max() <= max_iter_diff, (di.max(), np.where(di > max_iter_diff)[0])
    same = di == 0
    if same.any():  # same number of iterations: same arithmetic up to rounding
        assert rc[same].max() <= 1e-6
    assert np.all(alg.cost <= so["cost0"] * (1 + 1e-12))
    return float(same.mean())


@pytest.mark.parametrize("case,T,parallel", [("cart_pole", 100, False), ("acrobot", 80, True), ("double_parallel", 100, False)])
def test_ilqr_batch_converges_like_oracle(case, T, parallel):
    """Full solves of a batch of random starts, every trajectory (no percentile thresholds): see free_running_agreement."""
    rng = np.random.default_rng(5)
    n = len(CASES[case][4])
    B = 48
    x0 = np.array(CASES[case][4]) + rng.uniform(-0.3, 0.3, (B, n))
    alg = solver(case, T, x0=x0, parallel=parallel, maxIters=200)
    alg.run_optim()
    p = oracle_problem(case, S=4)
    so = oracle.ilqr_solve(p, oracle.ilqr_params(max_iters=200, parallel=parallel), x0, T, nthreads=8)
    free_running_agreement(alg, so, cost_tol=1e-5, max_iter_diff=8)


FULL_SIZE = {  # BASELINE.json configs[2] and configs[3] as bench.py builds them (per-GPU shard of rank 0)
    "c3": ("cart_pole", 16384, False, 500, 1000, [0, np.pi, 0, 0], [.5, .3, .5, .5]),
    "c4_double": ("double_parallel", 8192, True, 100, 4000, [0, np.pi, np.pi, 0, 0, 0], [.1] * 6),
    "c4_acrobot": ("acrobot", 8192, True, 100, 4000, [np.pi, 0, 0, 0], [.1] * 4),
}


@pytest.mark.parametrize("name", sorted(FULL_SIZE))
def test_full_size_ilqr_solves_match_oracle_on_a_sample(name):
    """The solves bench.py times, at their full sizes (16,384 swing-ups with the serial line search; 8,192 n = 6 / acrobot
    trajectories with `parallel=True`): (1) 256 sampled trajectories solved ALONE are bit-identical to their rows of the
    full batch -- costs, iteration counts, statuses, trajectories -- although the full batch runs the two-phase line search,
    the tail variant and CUDA-graph replay and the sample does not; (2) the oracle's free-running solves of the sample
    agree with them trajectory by trajectory (observed: C3 costs <= 1.3e-7, 178/256 identical iteration counts, max
    difference 6; C4 n = 6: all 256 iteration counts identical, costs 2e-15; acrobot: 254/256, costs 1e-12)."""
    case, B, par, iters, seed, nominal, spread = FULL_SIZE[name]
    rng = np.random.default_rng(seed)
    x0 = np.array(nominal) + rng.uniform(-1, 1, (B, len(nominal))) * np.array(spread)
    alg = solver(case, 100, x0=x0, maxIters=iters, parallel=par)
    alg.run_optim()
    sel = np.sort(np.random.default_rng(7).choice(B, 256, replace=False))
    sub = solver(case, 100, x0=x0[sel], maxIters=iters, parallel=par)
    sub.run_optim()
    assert np.array_equal(sub.cost, alg.cost[sel]) and np.array_equal(sub.iters, alg.iters[sel])
    assert np.array_equal(sub.status, alg.status[sel])
    assert np.array_equal(np.asarray(sub.xx), np.asarray(alg.xx)[sel]) and np.array_equal(np.asarray(sub.uu), np.asarray(alg.uu)[sel])
    so = oracle.ilqr_solve(oracle_problem(case, S=4), oracle.ilqr_params(max_iters=iters, parallel=par), x0[sel], 100, nthreads=16)
    same = free_running_agreement(sub, so, cost_tol=1e-6, max_iter_diff=8 if name == "c3" else 2)
    assert same >= (0.6 if name == "c3" else 0.97)


@pytest.mark.parametrize("case", ["cart_pole", "double_parallel"])
def test_fp32_ilqr_kernels_within_stated_tolerance(case):
    """The float instantiation of K2/K3/K4 (`dtype=torch.float32` environments) against the fp64 kernels on the same
    starts.  Stated tolerance: from identical nominal trajectories, ONE iteration -- gains K, k within 5e-5 of their largest
    entry, expected reductions within 5e-5, sane line-search candidate costs within 1e-5 (observed 3e-6 / 5e-6 / 4e-6 /
    9e-7: fp32 rounding over a 100-interval Riccati sweep and 400 RK4 substeps).  Whole solves: fp32 cannot resolve the
    reference's stopping test dcost < 1e-7 on costs of ~70, so it stops on the regularisation limit or the iteration cap
    instead, after more iterations; it must still reach the fp64 optimum: final costs within 1e-3 on >= 97 % of the
    trajectories (observed 1e-4 on 255/256) and never above the initial cost."""
    rng = np.random.default_rng(5)
    n = len(CASES[case][4])
    B, T = 256, 100
    x0 = np.array(CASES[case][4]) + rng.uniform(-0.3, 0.3, (B, n))
    a64 = solver(case, T, x0=x0, maxIters=200)
    a32 = solver(case, T, x0=x0, maxIters=200, env_kw={"dtype": torch.float32})
    assert a32._xbar.dtype == torch.float32 and a32._KKc.dtype == torch.float32
    assert rel(a32.cost, a64.cost) < 1e-5
    cost0 = a64.cost.copy()
    for a in (a64, a32):
        a._sel = a._select_params()
        a._rearm()
        a.iterate()
    K64, K32, k64, k32 = host(a64._KKc), host(a32._KKc), host(a64._kkc), host(a32._kkc)
    assert np.max(np.abs(K32 - K64) / np.max(np.abs(K64), axis=(1, 2, 3), keepdims=True)) < 5e-5
    assert np.max(np.abs(k32 - k64) / np.max(np.abs(k64), axis=(1, 2), keepdims=True)) < 5e-5
    d64, d32 = a64._bw["dV"].cpu().numpy(), a32._bw["dV"].double().cpu().numpy()
    assert np.max(np.abs(d32 - d64) / np.abs(d64)) < 5e-5
    c64, c32 = a64._fw["cost"].cpu().numpy(), a32._fw["cost"].double().cpu().numpy()
    sane = ~a64._fw["diverged"].cpu().numpy().astype(bool) & (c64 < 50 * (np.abs(cost0) + 1))
    assert np.array_equal(a32._fw["diverged"].cpu().numpy().astype(bool)[sane], np.zeros(int(sane.sum()), dtype=bool))
    assert np.max((np.abs(c32 - c64) / np.abs(c64))[sane]) < 1e-5
    a64 = solver(case, T, x0=x0, maxIters=200)
    a32 = solver(case, T, x0=x0, maxIters=200, env_kw={"dtype": torch.float32})
    a64.run_optim()
    a32.run_optim()
    rc = np.abs(a32.cost - a64.cost) / np.abs(a64.cost)
    assert np.mean(rc <= 1e-3) >= 0.97, np.sort(rc)[-10:]
    assert np.all(a32.cost <= cost0 * (1 + 1e-5))


def test_ilqr_per_trajectory_control_flow():
    """Trajectories that stop early must not be touched again, and a sharded batch must give exactly the
    bits of the full batch (pure data parallelism)."""
    rng = np.random.default_rng(6)
    B, T = 40, 60
    x0 = np.array([0, np.pi, 0, 0]) + rng.uniform(-0.3, 0.3, (B, 4))
    x0[3] = [0.01, 0.02, 0, 0]  # next to the upright equilibrium: stops early (at the optimum the reference's
    # loop may also leave through "no improvement": mu > mu_max)
    full = solver("cart_pole", T, x0=x0, maxIters=60)
    full.run_optim()
    assert full.iters[3] < full.iters.max() and full.status[3] in (pglib.ST_CONV_FUN, pglib.ST_CONV_GRAD, pglib.ST_MU_MAX)
    parts = []
    for sl in (slice(0, 17), slice(17, 40)):
        part = solver("cart_pole", T, x0=x0[sl], maxIters=60)
        part.run_optim()
        parts.append(part)
    for name in ("xx", "uu", "KK", "cost", "mu", "iters", "status"):
        assert np.array_equal(getattr(full, name), np.concatenate([getattr(p, name) for p in parts])), name


@pytest.mark.parametrize("case", ["cart_pole", "acrobot", "double_parallel"])
def test_fp32_rollout_within_stated_tolerance(case):
    """fp32 instantiation of the same kernels: 5e-3 relative on states over 100 control intervals, stated
    (north_star asks for a stated looser tolerance; measured values are recorded in DESIGN.md)."""
    rng = np.random.default_rng(7)
    n = len(CASES[case][4])
    B, T = 256, 100
    x0 = np.array(CASES[case][4]) + rng.uniform(-0.05, 0.05, (B, n))
    alg = solver(case, 10)
    umax = float(np.ravel(alg.environment.uMax)[0])
    U = rng.uniform(-umax, umax, (B, T, 1)) * 0.2
    dt = CASES[case][5]
    r64 = alg.lib.rollout(dev(x0), dev_U(U), S=4, dt=dt, history=True)
    r32 = alg.lib.rollout(dev(x0, torch.float32), dev_U(U, torch.float32), S=4, dt=dt, history=True)
    err = rel(r32["X"].double().cpu().numpy(), r64["X"].cpu().numpy())
    assert err < 5e-3, err
    assert rel(r32["cost"].double().cpu().numpy(), r64["cost"].cpu().numpy()) < 5e-3


def test_observation_kernel():
    env = make_env("cart_pole")
    L = env.library()
    rng = np.random.default_rng(8)
    xs = rng.uniform(-4, 4, (33, 4))
    o = host(L.observe(dev(xs)))
    want = np.stack([xs[:, 0], np.cos(xs[:, 1]), np.sin(xs[:, 1]), xs[:, 2], xs[:, 3]], axis=1)
    assert rel(o, want) < 1e-15
    assert rel(env.observe(xs), want) < 1e-15


# ---- properties at the benchmark's full size (BASELINE.json config 2) ---------------------------------
def test_full_size_rollout_properties():
    """65,536 envs x 500 steps: determinism, batch-permutation equivariance, horizon composition and
    history on/off -- all bit-exact -- plus an oracle check on a random subset."""
    B, T = 65536, 500
    alg = solver("cart_pole", 10)
    L = alg.lib
    gen = torch.Generator(device="cpu").manual_seed(0)
    x0 = torch.zeros((4, B), dtype=torch.float64)
    x0[0] = (torch.rand(B, generator=gen, dtype=torch.float64) - 0.5) * 0.1
    x0[1] = np.pi + (torch.rand(B, generator=gen, dtype=torch.float64) - 0.5) * 0.1
    U = (torch.rand((T, 1, B), generator=gen, dtype=torch.float64) * 2 - 1) * 10.0
    x0d, Ud = x0.cuda(), U.cuda()
    r = L.rollout(x0d, Ud, S=4, dt=0.02)
    r2 = L.rollout(x0d, Ud, S=4, dt=0.02)
    assert torch.equal(r["xN"], r2["xN"]) and torch.equal(r["cost"], r2["cost"])
    perm = torch.randperm(B, generator=gen).cuda()
    rp = L.rollout(x0d[:, perm].contiguous(), Ud[:, :, perm].contiguous(), S=4, dt=0.02)
    assert torch.equal(rp["xN"], r["xN"][:, perm]) and torch.equal(rp["cost"], r["cost"][perm])
    ra = L.rollout(x0d, Ud[:250].contiguous(), S=4, dt=0.02)
    rb = L.rollout(ra["xN"], Ud[250:].contiguous(), S=4, dt=0.02, t0=float(np.sum(np.full(250, 0.02))))
    assert torch.equal(rb["xN"], r["xN"])
    assert float(((ra["cost"] + rb["cost"]) - r["cost"]).abs().max()) < 1e-9 * float(r["cost"].abs().max())
    rh = L.rollout(x0d[:, :4096].contiguous(), Ud[:, :, :4096].contiguous(), S=4, dt=0.02, history=True)
    assert torch.equal(rh["xN"], r["xN"][:, :4096]) and torch.equal(rh["X"][-1], rh["xN"])
    idx = np.random.default_rng(9).choice(B, 64, replace=False)
    p = oracle_problem("cart_pole", S=4)
    ro = oracle.rollout(p, x0.numpy().T[idx], U=U.numpy()[:, :, idx].transpose(2, 0, 1), nthreads=8)
    assert rel(host(r["xN"])[idx], ro["xN"]) < 1e-8
    assert rel(r["cost"].cpu().numpy()[idx], ro["cost"]) < 1e-9


# ---- the reference-facing Python API -----------------------------------------------------------------
def test_environment_api_single_plant():
    """One plant, stepped like the reference's examples: shapes, history, tt, costs."""
    g = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "steps.npz"))
    env = make_env("cart_pole", x0=g["cart_pole/x0"])
    assert env.xDim == 4 and env.uDim == 1 and env.oDim == 5
    U = g["cart_pole/U"]
    costs = [env.step(U[k]) for k in range(10)]
    assert isinstance(costs[0], float) and np.shape(env.x) == (4,)
    assert env.history.shape == (11, 4) and len(env.tt) == 11
    assert rel(env.history, g["cart_pole/rk4x4/X"][:11]) < 1e-10
    assert rel(costs, g["cart_pole/rk4x4/c"][:10]) < 1e-10
    assert rel(env.x_, g["cart_pole/rk4x4/X"][9]) < 1e-10
    assert np.shape(env.o) == (5,) and env.terminated in (False, True)
    env.reset(g["cart_pole/x0"])
    env.rollout(U[:10])
    assert rel(env.history, g["cart_pole/rk4x4/X"][:11]) < 1e-10
    # numpy callables of the reference API
    assert rel(env.ode(0, [0.1, 0.7, -0.3, 1.1], [2.5]), [-0.3, 1.1, 2.5, 12.096269276790501]) < 1e-14
    assert env.A([0.1, 0.7, -0.3, 1.1], [2.5])[3, 1] == pytest.approx(8.708529569078923, rel=1e-13)
    assert env.B([0.1, 0.7, -0.3, 1.1], [2.5])[3, 0] == pytest.approx(1.1303497074638655, rel=1e-13)


def test_environment_api_batched_and_user_ode():
    rng = np.random.default_rng(10)
    x0 = np.array([np.pi, 0]) + rng.uniform(-0.1, 0.1, (7, 2))
    env = make_env("pendulum", x0=x0)
    U = rng.uniform(-3, 3, (7, 12, 1))
    cs = np.array([env.step(U[:, k]) for k in range(12)])
    assert cs.shape == (12, 7) and env.history.shape == (7, 13, 2)
    p = oracle_problem("pendulum", S=4)
    ro = oracle.rollout(p, x0, U=U)
    assert rel(env.history, ro["X"]) < 1e-10
    assert rel(cs.sum(0), ro["cost"]) < 1e-10
    assert not hasattr(env, "A")
    # a user-defined StateSpaceModel with numpy ufuncs in its ODE (like the reference's Pendulum.ode)
    from pygent_b200.examples import user_ode_env
    env2 = user_ode_env(x0)
    env2.rollout(U)
    assert rel(env2.history, ro["X"]) < 1e-10


def test_run_disk_loads_a_reference_controller(golden, tmp_path):
    """tests/golden/ilqr_disk/data/ holds a controller written by the UNMODIFIED reference's iLQR.save(); run_disk from
    another initial state must reproduce the reference's run_disk (oracle/make_golden_disk.py).  Then the round trip:
    this package's save() writes the same six files with the same shapes, and a batched solver reloads them."""
    import os
    import pickle
    import shutil
    from pygent_b200.algorithms.ilqr import iLQR
    from pygent_b200.examples import CASES, make_env
    g = golden("ilqr_disk")
    _, c_N, _ = CASES["cart_pole"][3]()
    src = os.path.join(os.path.dirname(__file__), "golden", "ilqr_disk", "data")
    path = str(tmp_path) + "/"
    shutil.copytree(src, path + "data")
    alg = iLQR(make_env("cart_pole", x0=list(g["x0b"])), 1.0, 0.02, fcost=c_N, constrained=False, path=path, printing=False,
               final_backpass=False, init=False)
    alg.run_disk(list(g["x0b"]))
    assert rel(alg.xx, g["same_xx"]) < 1e-9 and rel(alg.uu, g["same_uu"]) < 1e-9 and rel(alg.cost, g["same_cost"]) < 1e-9
    assert rel(alg.KK, g["same_KK"]) < 1e-12
    # zero-order-hold re-sampling to dt = 0.01: every stored gain is applied for two intervals
    env2 = type(alg.environment)(alg.environment._user_cost, list(g["x0b"]), 0.01)
    alg2 = iLQR(env2, 1.0, 0.01, fcost=c_N, constrained=False, path=path, printing=False, final_backpass=False, init=False)
    alg2.run_disk(list(g["x0b"]))
    assert alg2.KK.shape == (100, 1, 4) and np.array_equal(alg2.KK[0::2], g["same_KK"]) and np.array_equal(alg2.KK[1::2], g["same_KK"])
    assert np.isfinite(alg2.cost) and rel(alg2.xx[-1], g["same_xx"][-1]) < 0.2
    # round trip through this package's save(): same files, same shapes, loadable by a batch
    out = str(tmp_path) + "/mine/"
    alg3 = iLQR(make_env("cart_pole", x0=list(g["x0"])), 1.0, 0.02, fcost=c_N, constrained=False, path=out, printing=False,
                final_backpass=False)
    alg3.run_optim()
    assert rel(alg3.cost, g["cost"]) < 1e-7
    for f in ("K_.npy", "k.npy", "uu.npy", "xx.npy", "alpha.npy"):
        assert np.load(out + "data/" + f).shape == np.load(os.path.join(src, f)).shape, f
    with open(out + "data/time_info.p", "rb") as fh:
        assert len(pickle.load(fh)) == 51
    alg4 = iLQR(make_env("cart_pole", x0=list(g["x0b"])), 1.0, 0.02, fcost=c_N, constrained=False, path=out, printing=False,
                final_backpass=False, init=False)
    alg4.run_disk(list(g["x0b"]))
    assert rel(alg4.xx, g["same_xx"]) < 1e-5 and rel(alg4.cost, g["same_cost"]) < 1e-6


@pytest.mark.parametrize("case,B,T", [("cart_pole", 6000, 100), ("double_parallel", 5500, 50), ("cart_pole", 333, 45)])
def test_balanced_forward_is_bit_identical(case, B, T):
    """pg_ilqr_forward_balanced (persistent workers + ticket queue over group x alpha x time chunk) must return exactly
    the bits of pg_ilqr_forward: costs, flags, terminal states -- for the whole batch and for a compacted work list."""
    alg = solver(case, T, x0=None)
    n, m = alg.lib.n, alg.lib.m
    rng = np.random.default_rng(B)
    x0 = np.array(CASES[case][4]) + rng.uniform(-0.3, 0.3, (B, n))
    alg = solver(case, T, x0=x0)
    alg.backward_pass()                      # candidate gains along the u = 0 trajectory
    L, al = alg.lib, list(alg.alphas)
    args = (alg._x0, alg._xbar, alg._ubar, alg._KKc, alg._kkc)
    kw = dict(alphas=al, S=alg.substeps, dt=alg.dt, ulim=[float(np.ravel(alg.environment.uMax)[0])])
    a = L.ilqr_forward(*args, balanced=False, out={"xN": torch.zeros((len(al), n, B), dtype=torch.float64, device="cuda")}, **kw)
    b = L.ilqr_forward(*args, balanced=True, **kw)
    for k in ("cost", "diverged", "terminated", "xN"):
        assert torch.equal(a[k], b[k]), k
    L.check_queues()
    # compacted work list: only the listed trajectories are touched
    sel = torch.as_tensor(np.sort(rng.choice(B, B // 3, replace=False)), dtype=torch.int32, device="cuda")
    idx = torch.zeros(B, dtype=torch.int32, device="cuda")
    idx[: len(sel)] = sel
    cnt = torch.tensor([len(sel)], dtype=torch.int32, device="cuda")
    out = {"cost": torch.full_like(a["cost"], -1.0), "diverged": torch.full_like(a["diverged"], 7), "xN": torch.full_like(a["xN"], -1.0)}
    c = L.ilqr_forward(*args, balanced=True, work=(idx, cnt), out=out, **kw)
    mask = torch.zeros(B, dtype=torch.bool, device="cuda")
    mask[sel.long()] = True
    assert torch.equal(c["cost"][:, mask], a["cost"][:, mask]) and torch.equal(c["xN"][..., mask], a["xN"][..., mask])
    assert bool((c["cost"][:, ~mask] == -1.0).all()) and bool((c["diverged"][:, ~mask] == 7).all())


def test_triple_pendulum_ilqr_matches_oracle():
    """CartPoleTriple (n = 8, the largest state the register-resident kernels are instantiated for; environments.py:660-679):
    five iLQR iterations on a small batch against the C oracle -- initial cost, costs after each iteration's bookkeeping,
    trajectories and gains."""
    case, T, B = "triple", 40, 6
    rng = np.random.default_rng(11)
    x0 = np.array(CASES[case][4]) + rng.uniform(-0.05, 0.05, (B, 8))
    alg = solver(case, T, x0=x0, maxIters=5)
    p = oracle_problem(case, S=4)
    so = oracle.ilqr_solve(p, oracle.ilqr_params(max_iters=5), x0, T)
    assert rel(alg.cost, so["cost0"]) < 1e-9
    alg.run_optim()
    assert np.array_equal(alg.iters, so["iters"]) and np.array_equal(alg.status, so["status"])
    assert rel(alg.cost, so["cost"]) < 1e-7 and rel(alg.xx, so["xx"]) < 1e-6 and rel(alg.uu, so["uu"]) < 1e-5
    assert rel(alg.KK, so["KK"]) < 1e-5


@pytest.mark.parametrize("T", [15, 20, 21, 45, 100])
def test_ilqr_horizons_around_the_chunk_length_match_oracle(T):
    """The chunk-parallel commit restarts from snapshots every 20 control intervals: horizons below, at, just above and
    not a multiple of the chunk length (T <= 20 uses the single-rollout commit) against the C oracle, 8 iterations."""
    case, B = "cart_pole", 40
    rng = np.random.default_rng(T)
    x0 = np.array(CASES[case][4]) + rng.uniform(-0.3, 0.3, (B, 4))
    alg = solver(case, T, x0=x0, maxIters=8)
    so = oracle.ilqr_solve(oracle_problem(case, S=4), oracle.ilqr_params(max_iters=8), x0, T)
    alg.run_optim()
    assert np.array_equal(alg.iters, so["iters"]) and np.array_equal(alg.status, so["status"])
    assert rel(alg.cost, so["cost"]) < 1e-8 and rel(alg.xx, so["xx"]) < 1e-7 and rel(alg.uu, so["uu"]) < 1e-6
    assert rel(alg.KK, so["KK"]) < 1e-6 and rel(alg.kk[..., 0], so["kk"]) < 1e-6


@pytest.mark.parametrize("B,T", [(700, 60), (20000, 100)])
def test_two_phase_line_search_changes_nothing(B, T):
    """Serial line search evaluated in two launches (first 4 alphas for all, the other 7 only where nothing has succeeded
    yet): iterations, statuses, costs, trajectories and gains must be exactly those of evaluating all 11 for everybody."""
    rng = np.random.default_rng(B)
    x0 = np.array(CASES["cart_pole"][4]) + rng.uniform(-0.4, 0.4, (B, 4))
    res = []
    for nf in (0, 4):
        alg = solver("cart_pole", T, x0=x0, maxIters=12)
        alg.first_alphas = nf
        if B < 1000:
            alg._ls_small = 64   # two launches while more than 64 trajectories are live, one afterwards
        assert alg.B > alg._ls_small
        alg.run_optim()
        res.append((alg.iters.copy(), alg.status.copy(), np.array(alg.cost), np.array(alg.xx), np.array(alg.uu), np.array(alg.KK),
                    alg.current_alpha.copy() if hasattr(alg.current_alpha, "copy") else alg.current_alpha))
    for p, q in zip(*res):
        assert np.array_equal(p, q)


def test_tail_variant_changes_nothing():
    """Once the host has seen few live trajectories, iterate() runs the line search as one launch of the plain kernel
    (iLQR.tail_threshold): iterations, statuses, costs, trajectories and gains must be exactly those of never switching."""
    B, T = 700, 60
    rng = np.random.default_rng(77)
    x0 = np.array(CASES["cart_pole"][4]) + rng.uniform(-0.4, 0.4, (B, 4))
    res = []
    for thr in (0, 10 ** 9, 300):   # never / from the first count the host sees / when most have converged
        alg = solver("cart_pole", T, x0=x0, maxIters=80)
        alg.tail_threshold = thr
        alg.fused_tail = False
        alg._ls_small = 64
        alg.run_optim()
        assert alg._tail == (thr > 0) or thr == 300   # (300: only if the host saw the count while it was that low)
        res.append((alg.iters.copy(), alg.status.copy(), np.array(alg.cost), np.array(alg.xx), np.array(alg.uu), np.array(alg.KK),
                    np.array(alg.current_alpha)))
    for other in res[1:]:
        for p, q in zip(res[0], other):
            assert np.array_equal(p, q)


@pytest.mark.parametrize("case,kw", [("cart_pole", {}), ("cart_pole", {"constrained": True}), ("double_parallel", {"parallel": True}),
                                     ("pendulum", {"constrained": True}), ("car", {"constrained": True}), ("acrobot", {"regType": 1})])
def test_fused_tail_solver_changes_nothing(case, kw):
    """pg_ilqr_solve -- one launch in which a warp per trajectory loops backward sweep / all candidates side by side /
    decision / commit until its trajectory stops -- against the launch-per-step loop: iteration counts, statuses, costs,
    trajectories, gains, mu and the accepted alphas must be exactly equal, whether the fused solver takes over from the
    first count the host sees or only once most trajectories have stopped (serial and parallel search, box constraints with
    m = 1 and m = 2, analytic and finite-difference Jacobians, both regularisation types)."""
    B, T = 600, 60
    rng = np.random.default_rng(78)
    n = len(CASES[case][4])
    x0 = np.array(CASES[case][4]) + rng.uniform(-0.3, 0.3, (B, n))
    res = []
    for fused, thr in ((False, 0), (True, 10 ** 9), (True, 200)):
        alg = solver(case, T, x0=x0, maxIters=60, **kw)
        alg.tail_threshold, alg.fused_tail = thr, fused
        launches0 = alg.lib.launches
        alg.run_optim()
        res.append((alg.iters.copy(), alg.status.copy(), np.array(alg.cost), np.array(alg.xx), np.array(alg.uu), np.array(alg.KK),
                    np.array(alg.kk), np.array(alg.mu), np.array(alg.current_alpha)))
        if fused and thr > 10 ** 6 and alg.lib.m == 1:  # (m = 2 keeps the launch-per-step tail, see iLQR._solve_tail)
            assert alg.lib.launches - launches0 < 12 * alg.check_every  # one block of iterations, then ONE launch
    assert res[0][0].max() > 10
    for other in res[1:]:
        for p, q in zip(res[0], other):
            assert np.array_equal(p, q)


def test_rollout_host_random_blocking_and_pipelined():
    """`StateSpaceModel.rollout_host_random` (the call bench.py's `e2e` times): start states from pinned host memory,
    actions drawn in the kernel, results into pinned host memory.  The blocking call equals the device-resident rollout over
    the materialised action stream; calls submitted with `wait=False` (two alternating output buffers, step i + 1 in flight
    while step i is read) return, through their handles, exactly what the blocking calls return -- in every order of waiting
    the two-slot scheme allows."""
    from pygent_b200.examples import make_env
    B, T = 20480, 40
    rng = np.random.default_rng(12)
    x0 = np.array([0, np.pi, 0, 0]) + rng.uniform(-0.2, 0.2, (B, 4))
    env = make_env("cart_pole", x0=x0)
    L = env.library()
    pin = lambda *shape: torch.empty(shape, dtype=torch.float64).pin_memory()
    x0_pin = pin(4, B)
    x0_pin.copy_(torch.as_tensor(x0.T.copy()))
    ref = []
    for i in range(5):
        c, xn = pin(B), pin(4, B)
        r = env.rollout_host_random(x0_pin, T, 900 + i, cost_out=c, xN_out=xn)
        assert r["cost"] is c and r["xN"] is xn
        ref.append((c.clone(), xn.clone()))
    U = L.random_actions(B, T, 900, np.ravel(env.uMax).astype(float))
    d = L.rollout(dev(x0), U, S=env.substeps, dt=env.dt)
    assert torch.equal(d["cost"].cpu(), ref[0][0]) and torch.equal(d["xN"].cpu(), ref[0][1])
    bufs = [(pin(B), pin(4, B)) for _ in range(2)]
    pending, got = None, []
    for i in range(5):
        h = env.rollout_host_random(x0_pin, T, 900 + i, cost_out=bufs[i % 2][0], xN_out=bufs[i % 2][1], wait=False)
        if pending is not None:
            r = pending.wait()
            got.append((r["cost"].clone(), r["xN"].clone()))
        pending = h
    r = pending.wait()
    got.append((r["cost"].clone(), r["xN"].clone()))
    for (c0, x0_), (c1, x1) in zip(ref, got):
        assert torch.equal(c0, c1) and torch.equal(x0_, x1)


@pytest.mark.parametrize("case,rounds", [("cart_pole", 1), ("cart_pole", 7), ("double_parallel", 5), ("pendulum", 3), ("triple", 4),
                                         ("angular_pendulum", 6)])
def test_fused_tail_rounds_change_nothing(case, rounds):
    """The fused solver in rounds (pg_ilqr_solve.round_iters): every launch gives each running trajectory at most `rounds`
    iterations, writes the list of those still running and the next launch deals them to the warps anew (more trajectories
    than 4 x SMs, so that both ticket ranges and the re-dealing are exercised; lists at most that long are finished by the
    launch that finds them).  Whatever the round length, the solve must be the one of the launch-per-step loop, bit for bit."""
    B, T = 900, 50
    rng = np.random.default_rng(79)
    n = len(CASES[case][4])
    x0 = np.array(CASES[case][4]) + rng.uniform(-0.3, 0.3, (B, n)) * (0.2 if case == "triple" else 1.0)
    res = []
    for fused, r in ((False, 0), (True, rounds), (True, 10 ** 6)):
        alg = solver(case, T, x0=x0, maxIters=40)
        alg.tail_threshold, alg.fused_tail, alg.tail_round = 10 ** 9, fused, max(r, 1)
        alg.run_optim(host_results=False)
        res.append((alg.iters.copy(), alg.status.copy(), np.array(alg.cost), np.array(alg.xx), np.array(alg.uu), np.array(alg.KK),
                    np.array(alg.kk), np.array(alg.mu), np.array(alg.current_alpha)))
    assert res[0][0].max() > 10 and (res[0][0] < 40).any()
    for other in res[1:]:
        for p, q in zip(res[0], other):
            assert np.array_equal(p, q)


@pytest.mark.parametrize("case,T", [("cart_pole", 60), ("double_parallel", 60), ("double_serial", 40), ("pendulum", 40), ("car", 40), ("triple", 30),
                                    ("acrobot", 40)])
def test_terminal_value_on_device_matches_scipy(case, T):
    """pg_ilqr_terminal_value (Newton on the final cost + the doubling algorithm for the discrete Riccati equation, one
    thread per trajectory) against the reference's own route, scipy.optimize.minimize + scipy.linalg.solve_discrete_are
    (ilqr.py:246-260, `iLQR._terminal_value`), trajectory by trajectory; then the whole final backward pass: the gains of
    `final_on_device=True` against those of the host route."""
    rng = np.random.default_rng(9)
    n = len(CASES[case][4])
    B = 24
    x0 = np.array(CASES[case][4]) + rng.uniform(-0.2, 0.2, (B, n))
    alg = solver(case, T, x0=x0, maxIters=4, final_backpass=False)
    alg.run_optim()
    r = alg.lib.terminal_value(alg._xbar, alg.dt, float(alg.tt[-1]), lin_mode=alg.lin_mode)
    assert r is not None
    try:
        P_host = alg._terminal_value().cpu().numpy()                  # [n, n, B], per trajectory (B <= 32)
    except np.linalg.LinAlgError:
        # scipy finds no stabilising solution here (a mode on the unit circle the final cost does not see: the reference's
        # run_optim raises at this point).  The device route must then either say so per trajectory, or return a
        # symmetric solution of the Riccati equation: checked through the equation's residual.
        ok = r["ok"].cpu().numpy().astype(bool)
        if ok.any():
            P = np.moveaxis(r["P"].cpu().numpy(), -1, 0)[ok]
            xe = r["x_eq"].cpu().numpy().T[ok]
            u0 = np.zeros((len(xe), alg.uDim))
            Ad, Bd, _ = alg.linearization(xe, u0)
            Cxx, Cuu, Cxu, _, _ = alg.cost_lin(xe, u0, float(alg.tt[-1]))
            Ad, Bd, Cxx, Cuu, Cxu = (np.asarray(a).reshape((len(xe),) + np.asarray(a).shape[-2:]) for a in (Ad, Bd, Cxx, Cuu, Cxu))
            for i in range(len(xe)):
                G = Cuu[i] + Bd[i].T @ P[i] @ Bd[i]
                H = Bd[i].T @ P[i] @ Ad[i] + Cxu[i].T.reshape(G.shape[0], -1)
                res = Ad[i].T @ P[i] @ Ad[i] - P[i] - H.T @ np.linalg.solve(G, H) + Cxx[i]
                assert np.max(np.abs(res)) < 1e-7 * max(1.0, np.abs(P[i]).max()), np.max(np.abs(res))
                assert np.array_equal(P[i], P[i].T)
        return
    assert bool(r["ok"].all())
    P_dev = r["P"].cpu().numpy()
    scale = np.abs(P_host).max(axis=(0, 1), keepdims=True)
    # scipy's BFGS stops at gtol = 1e-5, Newton at 1e-14: the equilibria differ by ~1e-6, which moves P only where the
    # dynamics are nonlinear at x*; the Riccati solutions themselves agree to ~1e-10
    # (the pendulum chains on a cart amplify that difference most: 1.8e-6 observed for the serial one)
    tol = 1e-5 if case.startswith("double") else 1e-6
    assert np.max(np.abs(P_dev - P_host) / scale) < tol, np.max(np.abs(P_dev - P_host) / scale)
    assert np.max(np.abs(P_dev - np.swapaxes(P_dev, 0, 1))) == 0.0
    qf = np.array(CASES[case][3]()[2]["qf"], dtype=float)            # quadratic final costs: the minimiser is the origin in
    assert np.max(np.abs(r["x_eq"].cpu().numpy()[qf > 0])) < 1e-9      # every weighted direction (the others keep x_N's)
    gains = []
    for on_device in (False, True):
        a2 = solver(case, T, x0=x0, maxIters=4, final_backpass=True)
        a2.final_on_device = on_device
        a2.run_optim()
        assert a2.final_status == "ok"
        gains.append((np.array(a2.KK), np.array(a2.kk), np.array(a2.mu)))
    assert rel(gains[1][0], gains[0][0]) < tol and rel(gains[1][1], gains[0][1]) < tol
    assert np.array_equal(gains[0][2], gains[1][2])


def test_final_backward_pass_of_a_large_batch_is_per_trajectory():
    """B > 32 used to share ONE terminal value (trajectory 0's): now every trajectory gets its own, on the device.  Checked
    on a final cost whose minimiser depends on nothing but whose Riccati solution does not either -- so the per-trajectory
    values must all be equal to the host value of trajectory 0 -- and on sampled trajectories against scipy."""
    rng = np.random.default_rng(10)
    B, T = 4096, 50
    x0 = np.array(CASES["cart_pole"][4]) + rng.uniform(-0.3, 0.3, (B, 4))
    alg = solver("cart_pole", T, x0=x0, maxIters=6, final_backpass=True)
    assert alg.final_on_device is None
    alg.run_optim()
    assert alg.final_status == "ok"
    r = alg.lib.terminal_value(alg._xbar, alg.dt, float(alg.tt[-1]), lin_mode=alg.lin_mode)
    P = r["P"].cpu().numpy()
    assert bool(r["ok"].all())
    sub = solver("cart_pole", T, x0=x0[:8], maxIters=6, final_backpass=True)
    sub.final_on_device = False
    sub.run_optim()
    Ph = sub._terminal_value().cpu().numpy()
    assert np.max(np.abs(P[..., :8] - Ph) / np.abs(Ph).max()) < 1e-6
    assert rel(np.array(alg.KK)[:8], np.array(sub.KK)) < 1e-6
