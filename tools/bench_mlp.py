#!/usr/bin/env python
"""Learned-dynamics rollouts at BASELINE.json config-5 size: 32,768 trajectories x 100 Euler steps."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mlp_oracle as mo  # weights only (deterministic init), not timed
from pygent_b200 import mlp

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
ia = [False, True, False, False]
w, mom = mo.make_weights(6, 2), mo.make_moments(5, 1, 2)
dyn = mlp.MLPDynamics.from_arrays(4, 1, ia, w["W1"], w["b1"], w["W2"], w["b2"], w["W3"], w["b3"], **mom)
x0 = torch.zeros((4, B), dtype=torch.float64, device="cuda")
x0[1] = np.pi
U = (torch.rand((T, 1, B), dtype=torch.float64, device="cuda") * 2 - 1) * 8
flops = 2.0 * B * T * (6 * 500 + 500 * 500 + 500 * 2)
for prec, reps in (("bf16", 10), ("fp32", 2)):
    dyn.rollout_device(x0, 0.02, U=U, precision=prec)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        dyn.rollout_device(x0, 0.02, U=U, precision=prec)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print("%s: B=%d T=%d  %.3f ms  %.3e MLP env-steps/s  %.1f TFLOP/s (algorithmic 508 kFLOP/eval)" % (prec, B, T, ms, B * T / ms * 1e3, flops / ms / 1e9), flush=True)
