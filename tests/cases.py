"""Test-side view of the example cases: product definitions + the matching oracle problems."""
from pygent_b200.examples import CASES, ILQR_RUNS  # noqa: F401


def oracle_problem(case, S=4, integrator=0):
    import oracle
    _, _, system, costf, _, dt, box = CASES[case]
    _, _, cp = costf()
    n = oracle.system_dims(system)[0]
    lim = [box.get(i, 0.0) for i in range(n)]
    return oracle.make_problem(system, dt, S=S, integrator=integrator, term_lim=lim, **cp)


def rk45_tol(case):
    """Bound on |RK4 x 4 - the reference's native RK45| (relative, states over 40 random-action intervals).  1e-6 is
    north_star's figure; the stiffer torque/force-input pendulums get 1e-5 and the MarBot falling over under random
    inputs 5e-4: there the gap is the error of the reference's RK45 at its default rtol = 1e-3 (RK45 vs a converged
    RK4 x 32: 1.7e-4; RK4 x 4 vs RK4 x 32: 1.5e-6)."""
    if case in ("cart_pole", "cart_pole_tv", "pendulum", "acrobot", "double_parallel", "cart_pole_force", "car"):
        return 1e-6
    return 5e-4 if case == "marbot" else 1e-5
