"""Learned-dynamics rollouts and Jacobians: the reference-precision tensor-core path (f16x3) vs bf16 vs fp32 FFMA.
python tools/time_mlp_x3.py"""
import sys

import numpy as np
import torch

sys.path[:0] = [".", "tests"]
from oracle import mlp_oracle as mo
from pygent_b200 import mlp


def main():
    ia = [False, True, False, False]
    w, mom = mo.make_weights(6, 2), mo.make_moments(5, 1, 2)
    dyn = mlp.MLPDynamics.from_arrays(4, 1, ia, w["W1"], w["b1"], w["W2"], w["b2"], w["W3"], w["b3"], **mom)
    rng = np.random.default_rng(0)
    for R in (32768, 45056):
        T = 100
        x0 = torch.as_tensor(rng.uniform(-1, 1, (4, R)), device="cuda")
        U = torch.as_tensor(rng.uniform(-6, 6, (T, 1, R)), device="cuda")
        for prec in ("f16x3", "bf16"):
            for _ in range(2):
                X = dyn.rollout_device(x0, 0.02, U=U, precision=prec)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                X = dyn.rollout_device(x0, 0.02, U=U, precision=prec)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            print("rollout %6d rows x %d: %-6s %.3f ms  %.3e env-steps/s  %.0f TFLOP/s algorithmic (508 kflop/eval)" % (
                R, T, prec, ms, R * T / ms * 1e3, R * T * 508000.0 / ms * 1e3 / 1e12))
    T, B = 100, 4096
    xb = torch.as_tensor(rng.uniform(-1, 1, (T + 1, 4, B)), device="cuda")
    ub = torch.as_tensor(rng.uniform(-6, 6, (T, 1, B)), device="cuda")
    for prec in ("f16x3", "bf16"):
        for _ in range(2):
            dyn.linearize_device(xb, ub, 0.02, mode="analytic", precision=prec)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            dyn.linearize_device(xb, ub, 0.02, mode="analytic", precision=prec)
        b.record()
        torch.cuda.synchronize()
        print("jacobian chain %d points: %-6s %.3f ms" % (T * B, prec, a.elapsed_time(b) / 5))


if __name__ == "__main__":
    main()
