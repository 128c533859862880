#!/usr/bin/env python
"""Learned-dynamics iLQR at BASELINE.json config-5 size: B trajectories (32,768 / 8 GPUs = 4,096 per GPU), horizon 100,
MLP 6 -> 500 -> 500 -> 2.  Times the pieces of one iteration and a whole solve; prints Jacobian error statistics of the
tensor-core chain against the float64 chain rule.     python tools/bench_mlp_ilqr.py [B] [T] [iters]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mlp_oracle as mo  # deterministic weights + float64 chain rule (checker only)
from pygent_b200.algorithms.ilqr import iLQR
from pygent_b200.environments import LearnedStateSpaceModel
from pygent_b200.nn_models import NNDynamics

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dt = 0.02
IA = [False, True, False, False]


def c_k(x, u, mod):
    x1, x2, x3, x4 = x
    u1, = u
    return 10 * x1 ** 2 + 5 * x2 ** 2 + .01 * x3 ** 2 + .01 * x4 ** 2 + 0.1 * u1 ** 2


def c_N(x, mod):
    x1, x2, x3, x4 = x
    return 100 * x1 ** 2 + 100 * x2 ** 2 + 10 * x3 ** 2 + 10 * x4 ** 2


def make(x0, precision, lin):
    w, mom = mo.make_weights(6, 2), mo.make_moments(5, 1, 2)
    net = NNDynamics(4, 1, dxDim=2, oDim=5, xIsAngle=IA)
    with torch.no_grad():
        for i, l in enumerate((net.layer1, net.layer2, net.layer3), 1):
            l.weight.copy_(torch.from_numpy(w["W%d" % i]))
            l.bias.copy_(torch.from_numpy(w["b%d" % i]))
    for k, a in (("oMean", "o_mean"), ("oVar", "o_var"), ("uMean", "u_mean"), ("uVar", "u_var"), ("dxMean", "dx_mean"), ("dxVar", "dx_var")):
        setattr(net, k, torch.from_numpy(mom[a])[None])
    env = LearnedStateSpaceModel(net, c_k, x0, 1, dt, precision=precision, linearization=lin)
    env.uMax = np.array([8.0])
    return env, w, mom


def timed(fn, reps=3):
    fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


rng = np.random.default_rng(0)
x0 = np.array([0.3, np.pi - 0.4, 0.2, -0.5]) + rng.uniform(-.2, .2, (B, 4))
env, w, mom = make(x0, "bf16", "analytic")
alg = iLQR(env, T * dt, dt, fastForward=True, fcost=c_N, printing=False, maxIters=iters, final_backpass=False)
dyn = env.dynamics
EVAL = 2.0 * (6 * 500 + 500 * 500 + 500 * 2)  # flops of one network evaluation

# Jacobian error statistics on the initial trajectory (first 64 trajectories)
A, Bm = dyn.linearize_device(alg._xbar, alg._ubar, dt, mode="analytic", precision="bf16")
nb = min(B, 64)
xb = np.moveaxis(alg._xbar[:-1, :, :nb].cpu().numpy(), -1, 0).reshape(-1, 4)
ub = np.moveaxis(alg._ubar[:, :, :nb].cpu().numpy(), -1, 0).reshape(-1, 1)
Ae, Be = mo.exact_jacobian(w, mom, IA, xb, ub, dt)
Ag = np.moveaxis(A[..., :nb].cpu().numpy(), -1, 0).reshape(-1, 4, 4)
Bg = np.moveaxis(Bm[..., :nb].cpu().numpy(), -1, 0).reshape(-1, 4, 1)
for name, g, e in (("A", Ag[:, 2:], Ae[:, 2:]), ("B", Bg[:, 2:], Be[:, 2:])):
    err = np.abs(g - e) / np.abs(e).max()
    print("jacobian %s (tensor cores vs float64 chain rule): frobenius rel %.2e, |err|/scale median %.1e, 99%% %.1e, max %.1e, frac > 1.5e-2: %.4f"
          % (name, np.linalg.norm(g - e) / np.linalg.norm(e), np.median(err), np.percentile(err, 99), err.max(), np.mean(err > 1.5e-2)), flush=True)
Af, Bf = dyn.linearize_device(alg._xbar, alg._ubar, dt, mode="fd", precision="fp32")
Afn = np.moveaxis(Af[..., :nb].cpu().numpy(), -1, 0).reshape(-1, 4, 4)
print("jacobian A (fp32 central differences vs float64 chain rule): frobenius rel %.2e" % (np.linalg.norm(Afn[:, 2:] - Ae[:, 2:]) / np.linalg.norm(Ae[:, 2:])))

rows = B * T
t_an = timed(lambda: dyn.linearize_device(alg._xbar, alg._ubar, dt, mode="analytic", precision="bf16", out=alg._lin))
print("linearize analytic (bf16 tensor cores): %.3f ms for %d points = %.2e points/s; %.1f TFLOP/s counting 3 GEMM passes" % (
    t_an, rows, rows / t_an * 1e3, 3 * rows * 2.0 * 512 * 512 / t_an / 1e9), flush=True)
t_fd16 = timed(lambda: dyn.linearize_device(alg._xbar, alg._ubar, dt, mode="fd", precision="bf16", out=alg._lin), reps=1)
print("linearize central differences, bf16 evaluations (accuracy: useless, timing only): %.3f ms" % t_fd16, flush=True)
if rows <= 500000:
    t_fd = timed(lambda: dyn.linearize_device(alg._xbar, alg._ubar, dt, mode="fd", precision="fp32", out=alg._lin), reps=1)
    print("linearize central differences, fp32 FFMA (the reference's scheme): %.3f ms = %.1fx the analytic chain" % (t_fd, t_fd / t_an), flush=True)
alg._backward()
t_bw = timed(lambda: alg.lib.ilqr_backward_ext(alg._xbar, alg._ubar, alg._st["mu"], alg._lin["A"], alg._lin["B"], dt=dt, out=alg._bw))
print("backward (Riccati on given Jacobians): %.3f ms" % t_bw, flush=True)
al = list(alg.alphas)
t_fw = timed(lambda: dyn.forward_device(alg._x0, dt, al, (alg._KKc, alg._kkc, alg._xbar, alg._ubar), precision="bf16", out=alg._cand))
print("forward, 11 alphas (bf16 tensor cores): %.3f ms for %d rollouts x %d steps = %.2e MLP env-steps/s, %.1f TFLOP/s" % (
    t_fw, 11 * B, T, 11 * B * T / t_fw * 1e3, 11 * B * T * EVAL / t_fw / 1e9), flush=True)
t_c = timed(lambda: alg.lib.traj_cost(alg._cand["X"], alg._cand["U"], dt=dt, out=alg._fw))
print("candidate costs: %.3f ms" % t_c, flush=True)

for prec, lin in (("bf16", "analytic"),) + ((("fp32", "fd"),) if B <= 4096 else ()):
    env, _, _ = make(x0, prec, lin)
    alg = iLQR(env, T * dt, dt, fastForward=True, fcost=c_N, printing=False, maxIters=iters, final_backpass=False)
    c0 = float(np.mean(alg.cost))
    alg.run_optim()
    env.reset(x0)
    alg.reset()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    alg.run_optim()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    n_it = int(alg.iters.sum())
    print("%s/%s solve: B=%d T=%d, %d iterations in total (max %d) in %.1f ms = %.3e iLQR iterations/s; mean cost %.4f -> %.4f" % (
        prec, lin, B, T, n_it, int(alg.iters.max()), ms, n_it / ms * 1e3, c0, float(np.mean(alg.cost))), flush=True)
