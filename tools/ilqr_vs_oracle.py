"""Free-running agreement of GPU iLQR solves with the oracle at the BASELINE sizes: the full batch is solved on the GPU, a
sample of it by the oracle; prints how iteration counts / costs / statuses compare, and checks that the same sample solved
ALONE on the GPU is bit-identical to its rows in the full batch.   python tools/ilqr_vs_oracle.py [c3|c4|c4a] [sample]"""
import sys
import time

import numpy as np

sys.path[:0] = [".", "tests"]
import oracle
from cases import CASES, oracle_problem
from pygent_b200.examples import make_ilqr


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "c3"
    ns = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    if which == "c3":
        case, B, par, iters = "cart_pole", 16384, False, 500
        rng = np.random.default_rng(1000)
        x0 = np.array([0, np.pi, 0, 0]) + rng.uniform(-1, 1, (B, 4)) * np.array([.5, .3, .5, .5])
    elif which == "c4":
        case, B, par, iters = "double_parallel", 8192, True, 100
        rng = np.random.default_rng(4000)
        x0 = np.array([0, np.pi, np.pi, 0, 0, 0]) + rng.uniform(-0.1, 0.1, (B, 6))
    else:
        case, B, par, iters = "acrobot", 8192, True, 100
        rng = np.random.default_rng(4000)
        x0 = np.array([np.pi, 0, 0, 0]) + rng.uniform(-0.1, 0.1, (B, 4))
    T = 100
    alg = make_ilqr(case, T, x0=x0, maxIters=iters, parallel=par)
    t0 = time.time()
    alg.run_optim()
    print("GPU full batch: %.2f s, iterations total %d" % (time.time() - t0, int(alg.iters.sum())))
    sel = np.sort(np.random.default_rng(7).choice(B, ns, replace=False))
    sub = make_ilqr(case, T, x0=x0[sel], maxIters=iters, parallel=par)
    sub.run_optim()
    print("subset alone == rows of the full batch: cost %s iters %s status %s xx %s" % (
        np.array_equal(sub.cost, alg.cost[sel]), np.array_equal(sub.iters, alg.iters[sel]), np.array_equal(sub.status, alg.status[sel]),
        np.array_equal(np.asarray(sub.xx), np.asarray(alg.xx)[sel])))
    p = oracle_problem(case, S=4)
    t0 = time.time()
    so = oracle.ilqr_solve(p, oracle.ilqr_params(max_iters=iters, parallel=par), x0[sel], T, nthreads=16)
    print("oracle: %.2f s for %d solves" % (time.time() - t0, ns))
    gi, oi = alg.iters[sel], so["iters"]
    gc, oc = alg.cost[sel], so["cost"]
    gs, os_ = alg.status[sel], so["status"]
    rc = np.abs(gc - oc) / np.abs(oc)
    print("same iteration count: %d / %d; within 3: %d; max diff %d" % ((gi == oi).sum(), ns, (np.abs(gi - oi) <= 3).sum(), np.abs(gi - oi).max()))
    print("same status: %d / %d" % ((gs == os_).sum(), ns))
    for thr in (1e-9, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2):
        print("  cost rel diff <= %.0e: %d" % (thr, (rc <= thr).sum()))
    same = gi == oi
    print("where iteration counts agree: max cost rel diff %.2e" % (rc[same].max() if same.any() else 0))
    bad = np.where(rc > 1e-5)[0]
    for b in bad[:20]:
        print("  traj %d: iters %d/%d status %d/%d cost %.9f/%.9f" % (sel[b], gi[b], oi[b], gs[b], os_[b], gc[b], oc[b]))


if __name__ == "__main__":
    main()
