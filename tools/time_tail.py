"""Where a TAIL iteration of the 16,384-trajectory convergence run goes (few live trajectories): per-call CUDA-event
times of iLQR.iterate() after `skip` iterations.  Run on a B200."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from pygent_b200 import examples

skip = int(sys.argv[1]) if len(sys.argv) > 1 else 150
rng = np.random.default_rng(1000)
x0 = np.array([0, np.pi, 0, 0]) + rng.uniform(-1, 1, (16384, 4)) * np.array([.5, .3, .5, .5])
alg = examples.make_ilqr("cart_pole", 100, x0=x0, env_kw={"device": "cuda"}, maxIters=500)
alg._sel = alg._select_params()
alg._rearm()
for _ in range(skip):
    alg.iterate()
torch.cuda.synchronize()
print("live trajectories after", skip, "iterations:", int(alg._n_active.item()))
L = alg.lib
names = ["ilqr_backward", "linesearch_filter", "ilqr_forward", "ilqr_select", "ilqr_commit_chunked", "ilqr_commit"]
rec = []
for nm in names:
    f = getattr(L, nm)
    def wrap(*a, _f=f, _nm=nm, **k):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); r = _f(*a, **k); e1.record()
        rec.append((_nm, e0, e1))
        return r
    setattr(L, nm, wrap)
s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
n_it = 20
s0.record()
for _ in range(n_it):
    alg.iterate()
s1.record()
torch.cuda.synchronize()
tot = {}
for nm, a, b in rec:
    tot[nm] = tot.get(nm, 0.0) + a.elapsed_time(b) * 1e3 / n_it
print("per iteration us:", {k: round(v, 1) for k, v in tot.items()}, "sum", round(sum(tot.values()), 1), "iteration", round(s0.elapsed_time(s1) * 1e3 / n_it, 1))
