"""The 16,384-trajectory convergence run (BASELINE.json configs[2]) and small solves with and without the fused tail solver
(pg_ilqr_solve): device time of run_optim, best of 3.   python tools/time_fused_tail.py"""
import sys

import numpy as np
import torch

sys.path[:0] = [".", "tests"]
from pygent_b200.examples import make_ilqr


def timed(alg, xi, reps=3):
    best = 1e9
    for _ in range(reps):
        alg.environment.reset(xi)
        alg.reset()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        alg.run_optim(host_results=False)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    for B in (16384, 2048, 256):
        rng = np.random.default_rng(1000)
        xi = (np.array([0, np.pi, 0, 0]) + rng.uniform(-1, 1, (16384, 4)) * np.array([.5, .3, .5, .5]))[:B]
        out = []
        for fused in (False, True):
            alg = make_ilqr("cart_pole", 100, x0=xi, maxIters=500)
            alg.fused_tail = fused
            alg.run_optim()
            ms = timed(alg, xi)
            out.append((ms, int(alg.iters.sum()), int(alg.iters.max()), float(np.mean(alg.cost))))
        print("B = %5d: launch-per-step %.1f ms, fused tail %.1f ms  (iterations %d / %d, max %d, mean cost %.6f / %.6f)" % (
            B, out[0][0], out[1][0], out[0][1], out[1][1], out[1][2], out[0][3], out[1][3]))
        print("          per iteration of the longest trajectory: %.0f us -> %.0f us" % (out[0][0] * 1e3 / out[0][2], out[1][0] * 1e3 / out[1][2]))


def trace():
    import time
    rng = np.random.default_rng(1000)
    xi = np.array([0, np.pi, 0, 0]) + rng.uniform(-1, 1, (16384, 4)) * np.array([.5, .3, .5, .5])
    alg = make_ilqr("cart_pole", 100, x0=xi, maxIters=500)
    alg.run_optim()
    alg.environment.reset(xi)
    alg.reset()
    alg._trace = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    alg.run_optim()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print("live trajectories as the host saw them (iteration, live, ms since start):")
    print("  " + "  ".join("%d:%d@%.1f" % (i, l, (t - t0) * 1e3) for i, l, t in alg._trace))
    print("  total %.1f ms" % ((t1 - t0) * 1e3))
    if alg._round_trace:
        end = torch.cuda.Event(enable_timing=True)
        end.record()
        torch.cuda.synchronize()
        evs = [e for e, _ in alg._round_trace] + [end]
        print("rounds of the fused tail solver (live at start: ms):")
        print("  " + "  ".join("%d:%.2f" % (int(c), evs[i].elapsed_time(evs[i + 1])) for i, (_, c) in enumerate(alg._round_trace)))
    it = np.sort(alg.iters)
    print("iterations per trajectory: median %d, 90%% %d, 99%% %d, max %d; trajectories at max: %d" % (
        np.median(it), it[int(0.9 * len(it))], it[int(0.99 * len(it))], it[-1], int((it == it[-1]).sum())))


if __name__ == "__main__":
    trace()
    main()
