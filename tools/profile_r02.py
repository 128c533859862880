#!/usr/bin/env python
"""Small fixed workloads for the ncu captures of round 2 (one per kernel family; each finishes in seconds):
    python tools/profile_r02.py solve   # ilqr_solve_kernel: 1,024 cart-pole swing-ups, the fused tail solver from the first count on
    python tools/profile_r02.py mlp     # mlp_x3_kernel + mlp_jac_x3_kernel: 45,056-row line search, 409,600 linearisation points
    python tools/profile_r02.py train   # train_* kernels: 20 optimiser steps of batch 512"""
import os
import sys

import numpy as np
import torch

sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests")]
what = sys.argv[1] if len(sys.argv) > 1 else "solve"
rng = np.random.default_rng(0)
if what == "solve":
    from pygent_b200.examples import make_ilqr
    x0 = np.array([0, np.pi, 0, 0]) + rng.uniform(-1, 1, (1024, 4)) * np.array([.5, .3, .5, .5])
    alg = make_ilqr("cart_pole", 100, x0=x0, maxIters=120)
    alg.run_optim(host_results=False)
    torch.cuda.synchronize()
    print("solve ok", int(alg.iters.max()), float(np.mean(alg.cost)))
elif what == "mlp":
    from oracle import mlp_oracle as mo  # weights only
    from pygent_b200 import mlp
    w, mom = mo.make_weights(6, 2), mo.make_moments(5, 1, 2)
    dyn = mlp.MLPDynamics.from_arrays(4, 1, [False, True, False, False], w["W1"], w["b1"], w["W2"], w["b2"], w["W3"], w["b3"], **mom)
    R, T, B = 45056, 100, 4096
    x0 = torch.as_tensor(rng.uniform(-1, 1, (4, R)), device="cuda")
    U = torch.as_tensor(rng.uniform(-6, 6, (T, 1, R)), device="cuda")
    xb = torch.as_tensor(rng.uniform(-1, 1, (T + 1, 4, B)), device="cuda")
    ub = torch.as_tensor(rng.uniform(-6, 6, (T, 1, B)), device="cuda")
    for _ in range(3):
        X = dyn.rollout_device(x0, 0.02, U=U, precision="f16x3")
        dyn.linearize_device(xb, ub, 0.02, mode="analytic", precision="f16x3")
    torch.cuda.synchronize()
    print("mlp ok", float(X.abs().max()))
else:
    from pygent_b200.algorithms.mbrl import DynamicsTrainer
    from train_cases import dataset, fresh_net
    D = dataset()
    tk = DynamicsTrainer(fresh_net(), device="cuda", batch_size=512)
    tk.set_moments(D)
    tk.set_data(D)
    k = tk.backend.k
    idx = torch.as_tensor(np.arange(512, dtype=np.int32) % 1200, device="cuda")
    for _ in range(20):
        k.step(idx, 1e-3, 1e-3)
    torch.cuda.synchronize()
    print("train ok", float(k.bucket[-1]))
