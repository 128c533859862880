"""Times the serial line search evaluated in one launch (first_alphas=0) and in two (first_alphas=3):
full-batch iLQR iterations, the convergence run and the receding-horizon loop.  Run on a B200."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from pygent_b200 import examples
from pygent_b200.algorithms.nmpc import NMPC


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def ilqr(nf, B=16384, T=100):
    rng = np.random.default_rng(0)
    x0 = np.array([0.0, np.pi, 0.0, 0.0]) + rng.uniform(-0.4, 0.4, (B, 4))
    alg = examples.make_ilqr("cart_pole", T, x0=x0, env_kw={"device": "cuda"}, maxIters=500)
    alg.first_alphas = nf
    alg._sel = alg._select_params()
    alg._rearm()
    per = []
    for it in range(12):
        a = ev(); alg.iterate(); b = ev(); torch.cuda.synchronize()
        per.append(a.elapsed_time(b) * 1e3)
    alg2 = examples.make_ilqr("cart_pole", T, x0=x0, env_kw={"device": "cuda"}, maxIters=500)
    alg2.first_alphas = nf
    alg2.run_optim()
    alg2 = examples.make_ilqr("cart_pole", T, x0=x0, env_kw={"device": "cuda"}, maxIters=500)
    alg2.first_alphas = nf
    torch.cuda.synchronize(); t = time.time(); alg2.run_optim(); torch.cuda.synchronize()
    return per, (time.time() - t) * 1e3, int(np.sum(alg2.iters))


for nf in (0, 3, 2, 4):
    per, ms, its = ilqr(nf)
    print(f"first_alphas={nf}: iteration us {[round(p) for p in per]}  convergence run {ms:.1f} ms, {its} iterations, {its / ms * 1e3:.3e} it/s", flush=True)
