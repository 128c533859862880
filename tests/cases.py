"""Test-side view of the example cases: product definitions + the matching oracle problems."""
from pygent_b200.examples import CASES, ILQR_RUNS  # noqa: F401


def oracle_problem(case, S=4, integrator=0):
    import oracle
    _, _, system, costf, _, dt, box = CASES[case]
    _, _, cp = costf()
    n = oracle.system_dims(system)[0]
    lim = [box.get(i, 0.0) for i in range(n)]
    return oracle.make_problem(system, dt, S=S, integrator=integrator, term_lim=lim, **cp)
