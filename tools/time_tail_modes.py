"""Tail iterations of the 16,384-trajectory convergence run (few live trajectories) replayed from a CUDA graph of two,
for the line-search variants: split search on the queue kernel (default), one launch on the queue kernel, one launch on
the plain kernel.  PG_FWD_APC=k puts k alphas (warps) into one CTA of the plain kernel.  Run on a B200."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from pygent_b200 import examples, lib as pglib

skip = int(sys.argv[1]) if len(sys.argv) > 1 else 150
rng = np.random.default_rng(1000)
x0 = np.array([0, np.pi, 0, 0]) + rng.uniform(-1, 1, (16384, 4)) * np.array([.5, .3, .5, .5])


def run(mode):
    alg = examples.make_ilqr("cart_pole", 100, x0=x0, env_kw={"device": "cuda"}, maxIters=500)
    alg._sel = alg._select_params()
    alg._rearm()
    for _ in range(skip):
        alg.iterate()
    torch.cuda.synchronize()
    live = int(alg._n_active.item())
    orig = pglib.use_balanced_forward
    if mode != "split":
        alg.first_alphas = 0
    if mode == "plain":
        pglib.use_balanced_forward = lambda *a, **k: False
    try:
        g = alg._capture_two_iterations()
    finally:
        pglib.use_balanced_forward = orig
    for _ in range(3):
        g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    print("%-6s live %5d -> %5d: %.1f us per iteration" % (mode, live, int(alg._n_active.item()), a.elapsed_time(b) * 1e3 / 20), flush=True)


for mode in ("split", "single", "plain"):
    run(mode)
