#!/usr/bin/env python
"""Plain vs balanced rollout kernel on the same inputs: size of the difference (expected: none)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygent_b200.examples import make_ilqr  # noqa: E402

B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, int(sys.argv[2]) if len(sys.argv) > 2 else 500
L = make_ilqr("cart_pole", 10, x0=np.zeros((2, 4))).lib
gen = torch.Generator(device="cpu").manual_seed(B)
x0 = torch.zeros((4, B), dtype=torch.float64)
x0[0] = (torch.rand(B, generator=gen, dtype=torch.float64) - 0.5) * 2.0
x0[1] = np.pi + (torch.rand(B, generator=gen, dtype=torch.float64) - 0.5)
U = (torch.rand((T, 1, B), generator=gen, dtype=torch.float64) * 2 - 1) * 10.0
x0d, Ud = x0.cuda(), U.cuda()
for Tn in (1, 2, 25, 26, T):
    a = L.rollout(x0d, Ud[:Tn].contiguous(), S=4, dt=0.02, balanced=False)
    b = L.rollout(x0d, Ud[:Tn].contiguous(), S=4, dt=0.02, balanced=True)
    d = (a["xN"] - b["xN"]).abs()
    nz = (d > 0).any(0)
    print("T=%d: differing envs %d / %d, max abs diff %.3e, rel %.3e, first idx %s; cost diff %.3e" % (
        Tn, int(nz.sum()), B, float(d.max()), float((d / (a["xN"].abs() + 1)).max()), nz.nonzero()[:4].flatten().tolist(),
        float((a["cost"] - b["cost"]).abs().max())))
