"""Per-kernel times of one iLQR iteration on the triple pendulum on a cart (n = 8), for B trajectories.  Run on a B200."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from pygent_b200.examples import make_ilqr, CASES

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rng = np.random.default_rng(0)
x0 = np.array(CASES["triple"][4]) + rng.uniform(-0.05, 0.05, (B, 8))
alg = make_ilqr("triple", 100, x0=x0, maxIters=30)
alg._sel = alg._select_params()
alg._rearm()
for _ in range(3):
    alg.iterate()
torch.cuda.synchronize()
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
n_it = 5
for _ in range(n_it):
    alg.iterate()
torch.cuda.synchronize()
tot = {}
for nm, a, b in rec:
    tot[nm] = tot.get(nm, 0.0) + a.elapsed_time(b) * 1e3 / n_it
print("B =", B, "per iteration us:", {k: round(v, 1) for k, v in tot.items()}, "sum", round(sum(tot.values()), 1), "live", int(alg._n_active.item()))
