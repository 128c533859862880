#!/usr/bin/env python
"""Rollout kernel time vs batch size (scheduler quantisation experiment): tools/sweep_rollout.py B1 B2 ..."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygent_b200.examples import make_ilqr  # noqa: E402

alg = make_ilqr("cart_pole", 10, x0=np.zeros((2, 4)))
L = alg.lib
DTYPE = torch.float32 if os.environ.get("PG_DTYPE") == "f32" else torch.float64
for B in [int(b) for b in sys.argv[1:]]:
    x0 = torch.zeros((4, B), dtype=DTYPE, device="cuda")
    x0[1] = np.pi
    U = (torch.rand((500, 1, B), dtype=DTYPE, device="cuda") * 2 - 1) * 10
    for _ in range(3):
        L.rollout(x0, U, S=4, dt=0.02)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(20):
        L.rollout(x0, U, S=4, dt=0.02)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print("B=%7d warps/SMSP=%.2f  %.4f ms  %.3e env-steps/s" % (B, B / 32 / 592, ms, B * 500 / ms * 1e3), flush=True)
