#!/usr/bin/env python
"""Step-by-step comparison of the batched NMPC with the reference's closed-loop golden run."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygent_b200.algorithms.nmpc import NMPC
from pygent_b200.examples import CASES, make_env

g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", (sys.argv[2] if len(sys.argv) > 2 else "nmpc_cart_pole") + ".npz"))
dt, horizon = float(g["dt"]), float(g["horizon"])
_, c_N, _ = CASES["cart_pole"][3]()
x0 = list(g["x0"])
alg = NMPC(make_env("cart_pole", x0=x0), make_env("cart_pole", x0=x0), 1.0, dt, horizon, fcost=c_N, constrained=False,
           maxIters=500, step_iterations=1)
agent, opt = alg.agent, alg.agent.traj_optimizer
agent.load_plan(g["init_xx"], g["init_uu"], g["init_KK"], g["init_kk"], float(g["init_cost"]), float(g["init_alpha"]), env_x=g["init_env_x"])
x = torch.as_tensor(np.array(x0, dtype=float)[:, None].copy(), device="cuda")
L = alg.environment.library()
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 50):
    u = agent.take_action_device(x)
    acc = int(opt._st["accept"][0]); lt = int(opt._last_tried[0]); bw = int(opt._bw["success"][0])
    r = L.rollout(x, u[None], S=4, dt=dt, t0=i * dt)
    xn = r["xN"]
    print("step %2d alpha %.4f u %+.8f (gold %+.8f) acc %d (gold %d) last_tried %d bw %d mu %.3e plan cost %.6f (gold %.6f) planxN err %.2e  x err %.2e" % (
        i, float(opt._st["alpha_best"][0]), float(u[0, 0]), float(g["U"][i, 0]), acc, int(g["accepted"][i]), lt, bw, float(opt._st["mu"][0]), float(opt._st["cost"][0]),
        float(g["plan_costs"][i]), float(np.max(np.abs(opt._xbar[-1, :, 0].cpu().numpy() - g["plan_xN"][i]))),
        float(np.max(np.abs(xn[:, 0].cpu().numpy() - g["X"][i + 1])))), flush=True)
    x = xn
