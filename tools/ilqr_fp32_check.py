"""fp32 instantiation of the iLQR kernels against the fp64 one (same starts): first backward pass (K, k, sum dV), first
line search (11 candidate costs), whole solves (iterations, final cost).   python tools/ilqr_fp32_check.py [case]"""
import sys

import numpy as np
import torch

sys.path[:0] = [".", "tests"]
from cases import CASES
from pygent_b200.examples import make_ilqr


def host(t):
    return np.moveaxis(t.detach().to("cpu", torch.float64).numpy(), -1, 0)


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "cart_pole"
    rng = np.random.default_rng(5)
    n = len(CASES[case][4])
    B, T = 256, 100
    x0 = np.array(CASES[case][4]) + rng.uniform(-0.3, 0.3, (B, n))
    a64 = make_ilqr(case, T, x0=x0, maxIters=200)
    a32 = make_ilqr(case, T, x0=x0, maxIters=200, env_kw={"dtype": torch.float32})
    print("initial cost rel diff", np.max(np.abs(a32.cost - a64.cost) / np.abs(a64.cost)))
    for a in (a64, a32):
        a._sel = a._select_params()
        a._rearm()
        a.iterate()
    K64, K32 = host(a64._KKc), host(a32._KKc)
    k64, k32 = host(a64._kkc), host(a32._kkc)
    sc = np.max(np.abs(K64), axis=(1, 2, 3), keepdims=True)
    print("first backward pass: K max err / max|K| per trajectory: median %.2e max %.2e" % (np.median(np.max(np.abs(K32 - K64) / sc, axis=(1, 2, 3))), np.max(np.abs(K32 - K64) / sc)))
    sk = np.max(np.abs(k64), axis=(1, 2), keepdims=True)
    print("                     k: median %.2e max %.2e" % (np.median(np.max(np.abs(k32 - k64) / sk, axis=(1, 2))), np.max(np.abs(k32 - k64) / sk)))
    d64, d32 = a64._bw["dV"].cpu().numpy(), a32._bw["dV"].double().cpu().numpy()
    print("                     sum dV rel err max %.2e" % np.max(np.abs(d32 - d64) / (np.abs(d64) + 1e-12)))
    c64, c32 = a64._fw["cost"].cpu().numpy(), a32._fw["cost"].double().cpu().numpy()
    ok = ~(a64._fw["diverged"].cpu().numpy().astype(bool)) & (c64 < 50 * (np.abs(a64.cost) + 1))
    print("first line search: candidate cost rel err (sane candidates) median %.2e max %.2e" % (np.median((np.abs(c32 - c64) / np.abs(c64))[ok]), np.max((np.abs(c32 - c64) / np.abs(c64))[ok])))
    a64 = make_ilqr(case, T, x0=x0, maxIters=200)
    a32 = make_ilqr(case, T, x0=x0, maxIters=200, env_kw={"dtype": torch.float32})
    a64.run_optim()
    a32.run_optim()
    rc = np.abs(a32.cost - a64.cost) / np.abs(a64.cost)
    print("whole solves: status64 %s status32 %s" % (np.bincount(a64.status, minlength=5), np.bincount(a32.status, minlength=5)))
    print("  iterations 64: mean %.1f   32: mean %.1f" % (a64.iters.mean(), a32.iters.mean()))
    for thr in (1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1):
        print("  final cost rel diff <= %.0e: %d / %d" % (thr, (rc <= thr).sum(), B))
    print("  fp32 final cost <= fp64 initial cost everywhere:", bool(np.all(a32.cost <= a64.cost0 * (1 + 1e-5))) if hasattr(a64, "cost0") else "n/a")


if __name__ == "__main__":
    main()
