#!/usr/bin/env python
"""Small fixed workload for ncu: one rollout launch at benchmark size and a few iLQR iterations.
    python tools/profile_case.py [rollout|ilqr] [B] [iters]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygent_b200.examples import make_ilqr  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "rollout"
B = int(sys.argv[2]) if len(sys.argv) > 2 else (65536 if what == "rollout" else 16384)
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
rng = np.random.default_rng(0)
if what == "rollout":
    alg = make_ilqr("cart_pole", 10, x0=np.zeros((2, 4)))
    x0 = torch.zeros((4, B), dtype=torch.float64, device="cuda")
    x0[1] = np.pi
    U = (torch.rand((500, 1, B), dtype=torch.float64, device="cuda") * 2 - 1) * 10
    for _ in range(3):
        r = alg.lib.rollout(x0, U, S=4, dt=0.02)
    torch.cuda.synchronize()
    print("rollout ok", float(r["cost"].mean()))
else:
    case = os.environ.get("PG_CASE", "cart_pole")
    n = {"cart_pole": 4, "double_parallel": 6, "acrobot": 4}[case]
    nominal = {"cart_pole": [0, np.pi, 0, 0], "double_parallel": [0, np.pi, np.pi, 0, 0, 0], "acrobot": [np.pi, 0, 0, 0]}[case]
    x0 = np.array(nominal) + rng.uniform(-0.3, 0.3, (B, n))
    alg = make_ilqr(case, 100, x0=x0, maxIters=iters, parallel=bool(int(os.environ.get("PG_PARALLEL", "0"))))
    alg.run_optim()
    torch.cuda.synchronize()
    print("ilqr ok", alg.iterations, float(np.mean(alg.cost)))
