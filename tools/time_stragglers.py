"""Are the trajectories that run longest in the 16,384-trajectory solve (BASELINE.json configs[2]) slow per iteration, or
slowed by their neighbours?  Solves the full batch, then the 256 / 64 longest-running trajectories on their own, and 256
random ones.   python tools/time_stragglers.py"""
import sys

import numpy as np
import torch

sys.path[:0] = [".", "tests"]
from pygent_b200.examples import make_ilqr
from time_fused_tail import timed


def main():
    rng = np.random.default_rng(1000)
    xi = np.array([0, np.pi, 0, 0]) + rng.uniform(-1, 1, (16384, 4)) * np.array([.5, .3, .5, .5])
    alg = make_ilqr("cart_pole", 100, x0=xi, maxIters=500)
    alg.run_optim()
    iters = np.array(alg.iters)
    order = np.argsort(-iters, kind="stable")
    for name, sel in (("256 longest", order[:256]), ("64 longest", order[:64]), ("256 random", rng.permutation(16384)[:256])):
        sub = make_ilqr("cart_pole", 100, x0=xi[sel], maxIters=500)
        sub.run_optim()
        ms = timed(sub, xi[sel])
        it = np.array(sub.iters)
        print("%-12s: %.1f ms, iterations max %d mean %.0f -> %.0f us per iteration of the longest; same counts as in the full batch: %s" % (
            name, ms, it.max(), it.mean(), ms * 1e3 / it.max(), bool(np.array_equal(it, iters[sel]))))


if __name__ == "__main__":
    main()
