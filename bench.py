#!/usr/bin/env python
"""Benchmark of the hot path: batched cart-pole RK4 rollouts (BASELINE.json config 2) + batched iLQR.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the rollout kernel over the whole batch: 65,536 environments x 500 control
intervals x 4 RK4 substeps per GPU, random actions, fp64 (workload is per GPU: weak scaling; ranks share
nothing on the data path).  Prints ONE JSON line (rank 0):
  value      env-steps/s over all GPUs with inputs resident in HBM (CUDA events, max over ranks)
  e2e        the same through the public API (`env.rollout`) from pinned HOST buffers: H2D of x0 and U,
             kernel, D2H of the per-environment cost and final state, all inside the timed region
  roofline   FP64 FMA pipe: algorithmic flops per launch / mean launch duration vs the DFMA peak measured
             on this pool's B200 (profiles/fp_peak_b200.json), plus the HBM view of the same launch
  cpu_baseline  the CPU oracle (C port of the reference's step loop) on a bounded sample, all host threads
  ilqr       iLQR iterations/s on 16,384 cart-pole swing-ups per GPU (BASELINE.json config 3), to convergence
  fp32       the same rollout workload through the fp32 instantiation of the kernels (stated tolerance vs fp64: 5e-3
             relative on states over 100 control intervals), against the measured FFMA peak
  ilqr_c4    BASELINE.json configs[3]: 8,192 double-pendulum-on-a-cart (n = 6) and acrobot trajectories, horizon 100,
             parallel line search (`parallel=True`), 100 iterations at most
  nmpc       closed-loop control steps/s of 4,096 cart-pole plants under receding-horizon iLQR (NMPC/MPCAgent: one
             planner iteration + final backward pass + action + plan shift + plant step per control step, on the device)
  mlp_ilqr   iLQR iterations/s through the learned-MLP dynamics, 4,096 trajectories per GPU (config 5 = 32,768 over
             8 GPUs), tensor-core rollouts and Jacobian chain at the reference's fp32 precision (split-fp16 operands), 10
             iterations; `lower_precision_bf16`: the same on the bf16 tensor-core path (an explicitly approximate extra)
  train      optimiser steps/s of the dynamics training step (6 -> 500 -> 500 -> 2, batch 512 per GPU, Adagrad): hand-written
             tcgen05 forward / input-gradient / weight-gradient GEMMs + ONE NCCL all-reduce of the 255,003-float bucket per step
`--impl reference` times the reference's algorithm on the host cores (the C oracle port -- the reference
itself is pure Python and does not travel to the GPU box) on the same config/metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_ENVS, T_STEPS, SUBSTEPS, DT, U_MAX = 65536, 500, 4, 0.02, 10.0
FLOPS_PER_SUBSTEP = 228   # 4 f-evals (8 arith + 36 sincos) + 13n combine, n = 4, fp64
FLOPS_COST = 14           # quadratic running cost per control interval
ILQR_B, ILQR_T = 16384, 100
MLP_B, MLP_ITERS = 4096, 10
# dram__bytes_read.sum + dram__bytes_write.sum of one rollout_queue_kernel launch at the benchmark size, from the
# `ncu --set full` capture summarised in profiles/r01_rollout_queue_ncu_summary.txt (264.5 MB + 4.0 MB)
NCU_TRAFFIC_BYTES = 268.5e6
CONFIG = {"workload": "batched cart-pole RK4 rollouts, 65,536 envs x 500 steps per GPU, random actions "
                      "(BASELINE.json configs[1])",
          "system": "CartPole(linearized=True)", "envs_per_gpu": B_ENVS, "horizon": T_STEPS, "substeps": SUBSTEPS,
          "dt": DT, "integrator": "RK4", "cost": "examples/nmpc/cart_pole.py c_k",
          "l2_policy": "inputs larger than L2: U is 262 MB per GPU and is re-read from HBM every step"}


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: everything else that writes to file descriptor 1 (NCCL's version banner,
    library chatter) is sent to stderr; emit() writes the line to the real stdout, unbuffered."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    fd = _REAL_STDOUT if _REAL_STDOUT is not None else 1
    sys.stdout.flush()
    os.write(fd, line)


def algorithmic_flops_per_launch():
    return B_ENVS * T_STEPS * (SUBSTEPS * FLOPS_PER_SUBSTEP + FLOPS_COST)


def synthetic_inputs(seed, B=B_ENVS, T=T_STEPS):
    """x0 = [U(-.05,.05), pi+U(-.05,.05), 0, 0], u ~ U(-uMax, uMax); host, kernel layout."""
    import torch
    gen = torch.Generator(device="cpu").manual_seed(seed)
    x0 = torch.zeros((4, B), dtype=torch.float64)
    x0[0] = (torch.rand(B, generator=gen, dtype=torch.float64) - 0.5) * 0.1
    x0[1] = np.pi + (torch.rand(B, generator=gen, dtype=torch.float64) - 0.5) * 0.1
    U = (torch.rand((T, 1, B), generator=gen, dtype=torch.float64) * 2 - 1) * U_MAX
    return x0, U


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [c.strip() for c in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0.5 * max(mx)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def fp64_peak():
    p = os.path.join(ROOT, "profiles", "fp_peak_b200.json")
    if os.path.isfile(p):
        with open(p) as fh:
            return float(json.load(fh)["fp64_fma_tflops"]), "measured DFMA microbenchmark on this pool's B200 (profiles/fp_peak_b200.json, tools/fp_peak.cu)"
    return 37.2, "nominal 148 SM x 64 lanes x 2 x 1.965 GHz (no measurement committed)"


def fp32_peak():
    p = os.path.join(ROOT, "profiles", "fp_peak_b200.json")
    if os.path.isfile(p):
        with open(p) as fh:
            return float(json.load(fh)["fp32_fma_tflops"])
    return 74.4


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_rollout_baseline(n_envs, threads, seed=0):
    """The oracle's step loop (RK4 x 4) on the first `n_envs` environments of the workload."""
    import oracle
    from pygent_b200.examples import CASES
    _, _, cp = CASES["cart_pole"][3]()
    p = oracle.make_problem("cart_pole_lin", DT, S=SUBSTEPS, term_lim=[1.2, 0, 4.0, 0], **cp)
    x0, U = synthetic_inputs(seed, B=n_envs)
    x0h = np.ascontiguousarray(x0.numpy().T)
    Uh = np.ascontiguousarray(U.numpy().transpose(2, 0, 1))
    t0 = time.perf_counter()
    oracle.rollout(p, x0h, U=Uh, history=False, nthreads=threads)
    dt = time.perf_counter() - t0
    return n_envs * T_STEPS / dt, dt


def cpu_ilqr_baseline(n_traj, threads, seed=1000):
    """M2 on the host: the C oracle's `run_optim` (line-by-line port of ilqr.py:399-485, RK4 x 4) on the first `n_traj` of
    the C3 starts bench.py's `ilqr` leg solves on the GPU (rank 0's shard), OpenMP over trajectories.  Returns (it/s, s, its)."""
    import oracle
    from pygent_b200.examples import CASES
    _, _, cp = CASES["cart_pole"][3]()
    p = oracle.make_problem("cart_pole_lin", DT, S=SUBSTEPS, term_lim=[1.2, 0, 4.0, 0], **cp)
    rng = np.random.default_rng(seed)
    xi = np.array([0, np.pi, 0, 0]) + rng.uniform(-1, 1, (ILQR_B, 4)) * np.array([.5, .3, .5, .5])
    t0 = time.perf_counter()
    so = oracle.ilqr_solve(p, oracle.ilqr_params(max_iters=500), xi[:n_traj], ILQR_T, nthreads=threads)
    secs = time.perf_counter() - t0
    its = int(so["iters"].sum())
    return its / secs, secs, its


def cpu_ilqr_report(threads):
    """`ilqr.cpu_baseline` of the bench line: the oracle port on all host threads and on one, and -- when oracle/_ref (the
    pip-installed unmodified reference, oracle/install_ref.py) is present -- the reference's own Python path on C1."""
    n = 512
    v, secs, its = cpu_ilqr_baseline(n, threads)
    v1, secs1, its1 = cpu_ilqr_baseline(32, 1)
    out = {"value": v, "unit": "iLQR iterations/s", "cores": threads, "kind": "port", "single_core": v1,
           "sample": "first %d of the 16,384 C3 swing-ups to convergence (%d iterations, %.1f s), C oracle, OpenMP over %d threads; "
                     "single_core: first 32 on one thread (%d iterations, %.1f s)" % (n, its, secs, threads, its1, secs1)}
    try:
        from oracle import ref_bench
        if ref_bench.available():
            r = ref_bench.run_c1()
            out["reference_python"] = {"value": r["iterations"] / r["seconds"], "unit": "iLQR iterations/s", "cores": 1, "kind": "reference",
                                       "iterations": r["iterations"], "final_cost": r["cost"], "seconds": r["seconds"],
                                       "expected": {"iterations": ref_bench.C1_ITERATIONS, "final_cost": ref_bench.C1_COST},
                                       "matches_known_answer": bool(r["iterations"] == ref_bench.C1_ITERATIONS and abs(r["cost"] - ref_bench.C1_COST) < 1e-9),
                                       "sample": "BASELINE.json configs[0]: the unmodified reference (oracle/_ref), one cart-pole swing-up, "
                                                 "horizon 100, RK45, lambdified sympy, one Python thread"}
        else:
            out["reference_python"] = {"unavailable": "oracle/_ref not installed (python -m oracle.install_ref needs /root/reference)"}
    except Exception as e:  # the baseline must never take the bench line down
        out["reference_python"] = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    return out


def run_reference(args, rank):
    """Reference arm: the reference's algorithm on the host cores (oracle port), same metric and config."""
    if rank != 0:
        return
    import oracle
    oracle.build()
    threads = os.cpu_count() or 1
    n = 8192
    for _ in range(args.warmup):
        cpu_rollout_baseline(n, threads)
    el = 0.0
    for _ in range(args.steps):
        el += cpu_rollout_baseline(n, threads)[1]
    value = n * T_STEPS * args.steps / el
    sample = "%d of the 65,536 envs x 500 steps per step (RK4 x 4), OpenMP over all host threads" % n
    ilqr_cpu = cpu_ilqr_report(threads)
    emit({
        "impl": "reference", "metric": "rk4_env_steps_per_s", "value": value, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": CONFIG,
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ilqr": {"metric": "ilqr_iterations_per_s", "value": ilqr_cpu["value"], "unit": "iLQR iterations/s", "cpu_baseline": ilqr_cpu}})



def ilqr_full_batch_roofline(xi, dev, n_iter=8):
    """FP64-pipe roofline of iLQR while the whole batch is live: the first n_iter iterations of a fresh solve, timed with
    CUDA events.  Algorithmic flops per iteration and trajectory: 570 per backward step + per alpha the serial
    line search of the reference TRIES (last alpha index + 1) a closed-loop rollout of T x (S x 228 + 30)."""
    import torch
    from pygent_b200.examples import make_ilqr
    solver = make_ilqr("cart_pole", ILQR_T, x0=xi, env_kw={"device": dev}, maxIters=500)
    tried_flops = ILQR_T * (SUBSTEPS * FLOPS_PER_SUBSTEP + 30)
    res = None
    for rep in range(2):  # the first repetition warms up
        solver.environment.reset(xi)
        solver.reset()
        solver._sel = solver._select_params()
        solver._rearm()
        tried, live = [], []
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(n_iter):
            live.append(solver._st["active"].clone())
            solver.iterate()
            tried.append(solver._last_tried.clone())
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        n_live = sum(int(l.sum()) for l in live)
        n_tried = sum(int(((t + 1) * l.to(torch.int32)).sum()) for t, l in zip(tried, live))
        flops = n_live * ILQR_T * 570 + n_tried * tried_flops
        res = {"iterations": n_live, "us_per_iteration": ms * 1e3 / n_iter, "value": n_live / (ms * 1e-3),
               "unit": "iLQR iterations/s", "alphas_tried_mean": n_tried / max(n_live, 1),
               "alphas_evaluated": "first 4 of 11 for all, the other 7 where none of those passed",
               "achieved": flops / (ms * 1e-3) / 1e12, "unit_roofline": "TFLOP/s", "bound": "fp64"}
    return res


def nmpc_bench(dev, rank, world, barrier, plants=4096, steps=102, horizon=1.0):
    """examples/nmpc/cart_pole.py for a batch: every plant has its own MPCAgent state; the planner's terminal value is
    solved once on the host (cache_terminal_value) so the control loop never leaves the device: after two eager control
    steps the step is captured into a CUDA graph and replayed (NMPC.run), which keeps the host out of the timed loop."""
    import torch
    import torch.distributed as dist
    from pygent_b200.algorithms.nmpc import NMPC
    from pygent_b200.examples import CASES, make_env
    _, c_N, _ = CASES["cart_pole"][3]()
    rng = np.random.default_rng(3000 + rank)
    x0 = np.array([0, 0, 0, 0]) + rng.uniform(-1, 1, (plants, 4)) * np.array([.2, .25, .2, .2])
    alg = NMPC(make_env("cart_pole", x0=x0, device=dev), make_env("cart_pole", x0=x0, device=dev), steps * DT, DT, horizon,
               fcost=c_N, constrained=False, maxIters=100, step_iterations=1, cache_terminal_value=True)
    opt = alg.agent.traj_optimizer
    alg.run(steps=steps, printing=False)  # warm-up; captures the 10-control-step CUDA graph that the timed run replays
    alg.agent.init_optim(x0)
    barrier()
    l0 = opt.lib.launches + alg.environment.library().launches
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    cost = alg.run(steps=steps, printing=False)
    b.record()
    barrier()
    ms = a.elapsed_time(b)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    return {"metric": "nmpc_control_steps_per_s", "value": world * plants * steps / (ms * 1e-3), "unit": "plant control steps/s",
            "plants_per_gpu": plants, "control_steps": steps, "horizon_steps": int(horizon / DT), "ms": ms,
            "ms_per_control_step": ms / steps, "gpu_launches": opt.lib.launches + alg.environment.library().launches - l0,
            "mean_closed_loop_cost": float(np.mean(np.sum(cost, axis=-1))),
            "stabilised_frac": float(np.mean(np.abs(alg.environment.history[:, -1, 1]) < 0.2))}


def learned_ilqr_bench(dev, rank, world, barrier, precision="f16x3"):
    """BASELINE.json configs[4] per GPU: iLQR over NNDynamics(4, 1, oDim=5, dxDim=2) with default-initialised weights
    (torch.manual_seed(0)), moments 0/1, Euler, horizon 100, costs of examples/mbrl/cart_pole_hpc.py:24-33."""
    import torch
    import torch.distributed as dist
    from pygent_b200.algorithms.ilqr import iLQR
    from pygent_b200.environments import LearnedStateSpaceModel
    from pygent_b200.nn_models import NNDynamics

    def c_k(x, u, mod):
        x1, x2, x3, x4 = x
        u1, = u
        return 10 * x1 ** 2 + 5 * x2 ** 2 + .01 * x3 ** 2 + .01 * x4 ** 2 + 0.1 * u1 ** 2

    def c_N(x, mod):
        x1, x2, x3, x4 = x
        return 100 * x1 ** 2 + 100 * x2 ** 2 + 10 * x3 ** 2 + 10 * x4 ** 2

    torch.manual_seed(0)
    net = NNDynamics(4, 1, dxDim=2, oDim=5, xIsAngle=[False, True, False, False])
    rng = np.random.default_rng(2000 + rank)
    x0 = np.array([0, np.pi, 0, 0]) + rng.uniform(-1, 1, (MLP_B, 4)) * np.array([.5, .3, .5, .5])
    env = LearnedStateSpaceModel(net, c_k, x0, 1, DT, precision=precision, linearization="analytic", device=dev)
    env.uMax = np.array([U_MAX])
    alg = iLQR(env, ILQR_T * DT, DT, fastForward=True, fcost=c_N, printing=False, maxIters=MLP_ITERS, final_backpass=False)
    c0 = float(np.mean(alg.cost))
    alg.run_optim()  # warm-up solve (same work)
    env.reset(x0)
    alg.reset()
    barrier()
    l0 = alg.lib.launches + env.dynamics.launches
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    alg.run_optim()
    b.record()
    barrier()
    ms, total = a.elapsed_time(b), int(alg.iters.sum())
    if world > 1:
        t = torch.tensor([ms, -float(total)], dtype=torch.float64, device=dev)
        tm, ts = t.clone(), t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        ms, total = float(tm[0]), int(-ts[1])
    # algorithmic work: one network evaluation = 2 (6 x 500 + 500 x 500 + 500 x 2) = 508,000 flops; per iteration and
    # trajectory T x (11 line-search rollouts + the Jacobian chain: 1 forward + dx_dim = 2 transposed 500 x 500 products)
    evals = total * ILQR_T * (11 + 3)
    tf = evals * 508000.0 / (ms * 1e-3) / 1e12 / world
    peak = 1384.0
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peak = float(json.load(fh)["bf16_tflops_sustained"])
    except (OSError, KeyError, ValueError):
        pass
    mmas = 3 if precision == "f16x3" else 1
    return {"metric": "mlp_ilqr_iterations_per_s", "value": total / (ms * 1e-3), "unit": "iLQR iterations/s",
            "trajectories_per_gpu": MLP_B, "horizon": ILQR_T, "iterations_total": total, "ms": ms, "precision": precision,
            "network": {"f16x3": "6->500->500->2; layer 2 on tcgen05 with fp16 hi/lo split operands (Ah Wh + Al Wh + Ah Wl), fp32 accumulate: "
                                 "the reference's fp32 precision (stated 1e-5 per evaluation, tests/test_gpu_mlp.py)",
                        "bf16": "6->500->500->2; layer 2 on tcgen05 with bf16 operands, fp32 accumulate: LOWER precision than the reference "
                                "(2e-2 per evaluation) -- reported as an extra, not as the config-5 result"}[precision],
            "linearization": "analytic chain on tensor cores",
            "tflops": tf, "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                                       "peak_source": "measured sustained bf16 cuBLAS GEMM (MEASURED_PEAKS.json)", "mma_per_product": mmas,
                                       "frac_executed": mmas * tf / peak,
                                       "note": "algorithmic flops of the 6->500->500->2 network (508,000 per evaluation, unpadded); the split-fp16 "
                                               "path executes 3 tensor-core MMAs per algorithmic product"},
            "gpu_launches": alg.lib.launches + env.dynamics.launches - l0,
            "mean_cost_before": c0, "mean_cost_after": float(np.mean(alg.cost))}


TRAIN_ROWS = 512   # the reference's batch size (mbrl.py:50): per GPU here (weak scaling of the data-parallel step)


def synthetic_transitions_np(n, seed):
    """Cart-pole-like transitions with the column layout MBRL stores (mbrl.py:160-167): float32 arrays."""
    rng = np.random.RandomState(seed)
    x_ = np.array([0, np.pi, 0, 0]) + rng.uniform(-1, 1, (n, 4)) * np.array([.5, 1.5, 2., 4.])
    u = rng.uniform(-8, 8, (n, 1))
    f = np.stack([x_[:, 2], x_[:, 3], u[:, 0], 1.5 * u[:, 0] * np.cos(x_[:, 1]) + 14.5 * np.sin(x_[:, 1])], axis=1)
    x = x_ + 0.02 * f + rng.normal(0, 1e-3, (n, 4))
    o_ = np.stack([x_[:, 0], np.cos(x_[:, 1]), np.sin(x_[:, 1]), x_[:, 2], x_[:, 3]], axis=1)
    return {k: v.astype(np.float32) for k, v in (("x_", x_), ("u", u), ("x", x), ("o_", o_))}


def train_bench(dev, rank, world, barrier, steps=300, with_cpu=True):
    """The dynamics training step: NNDynamics(4, 1, oDim=5, dxDim=2), batch 512 rows per GPU (global batch
    512 x N, every rank back-propagates its strided slice), Adagrad(lr 1e-3, weight decay 1e-3), one all-reduce of the flat
    gradient bucket per step (NCCL, on the compute stream).  Device-timed with CUDA events, max over ranks."""
    import torch
    import torch.distributed as dist
    from pygent_b200.algorithms.mbrl import DynamicsTrainer
    from pygent_b200.nn_models import NNDynamics
    torch.manual_seed(0)
    net = NNDynamics(4, 1, dxDim=2, oDim=5, xIsAngle=[False, True, False, False])
    gb = TRAIN_ROWS * world
    tr = DynamicsTrainer(net, dyn_lr=1e-3, weight_decay=1e-3, batch_size=gb, device=dev)
    col = synthetic_transitions_np(8 * gb, seed=11)

    class _D:  # the DataSet methods the trainer uses
        rng = np.random.RandomState(5)

        def columns(self):
            return col

        def batch_indices(self, n, shuffle=True):
            idx = np.arange(8 * gb)
            if shuffle:
                self.rng.shuffle(idx)
            return [idx[i:i + n] for i in range(0, len(idx), n)]
    tr.set_moments(_D())
    tr.set_data(col)
    batches = [np.arange(i * gb, (i + 1) * gb) for i in range(8)]
    k = tr.backend.k
    idx_dev = [torch.as_tensor(b[rank::world].astype(np.int32), device=dev) for b in batches]

    def step(i):  # DynamicsTrainer.train_step with the index upload hoisted out of the loop
        k.grad(idx_dev[i % 8], gb)
        if world > 1:
            dist.all_reduce(k.bucket, op=dist.ReduceOp.SUM)
        k.update(1e-3, 1e-3)
    loss0 = None
    for i in range(10):
        step(i)
        if loss0 is None:
            loss0 = float(k.bucket[-1])
    barrier()
    l0 = k.launches
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        step(i)
    b.record()
    barrier()
    ms = a.elapsed_time(b)
    launches = k.launches - l0
    loss1 = float(k.bucket[-1])
    # the collective alone, back to back on the compute stream
    ar_us = None
    if world > 1:
        for _ in range(5):
            dist.all_reduce(k.bucket, op=dist.ReduceOp.SUM)
        barrier()
        a.record()
        for _ in range(100):
            dist.all_reduce(k.bucket, op=dist.ReduceOp.SUM)
        b.record()
        barrier()
        ar_us = a.elapsed_time(b) * 10.0
        t = torch.tensor([ms, ar_us], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ar_us = float(t[0]), float(t[1])
    # the step without the collective (grad + update only), same loop
    barrier()
    a.record()
    for i in range(steps):
        k.grad(idx_dev[i % 8], gb)
        k.update(1e-3, 1e-3)
    b.record()
    barrier()
    ms_local = a.elapsed_time(b)
    # ---- the same step with the fused NVLink peer-memory exchange (pg_mlp_train_update_p2p) instead of NCCL + update ----
    p2p = None
    if world > 1:
        try:
            k.enable_p2p()
            for i in range(10):
                k.grad(idx_dev[i % 8], gb, bucket=k.p2p_bucket())
                k.update_p2p(1e-3, 1e-3)
            barrier()
            a.record()
            for i in range(steps):
                k.grad(idx_dev[i % 8], gb, bucket=k.p2p_bucket())
                k.update_p2p(1e-3, 1e-3)
            b.record()
            barrier()
            t = torch.tensor([a.elapsed_time(b), float(k.p2p_status())], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            p2p = {"value": steps / (float(t[0]) * 1e-3), "unit": "optimiser steps/s", "us_per_step": float(t[0]) * 1e3 / steps,
                   "gave_up_waiting": int(t[1]), "loss_last": float(k.loss_dev[0]),
                   "exchange": "each rank writes its gradient into its own NVLink-mapped block; ONE kernel waits for all ranks, sums the "
                               "world's buckets by peer loads in rank order and applies Adagrad (no NCCL call in the step)"}
        except Exception as e:  # noqa: BLE001 -- peer mapping unavailable on this node: the NCCL figures stand
            p2p = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    # ---- the public call: DynamicsTrainer.training(data set) -- moments, upload, `epochs` shuffled passes of 8 batches; a single
    # process replays every epoch from one CUDA graph (host work per epoch: one shuffle, one index upload, one loss read-back)
    api = None
    if world == 1:
        tr.training_epochs = 5
        tr.training(_D())   # captures the epoch graph
        tr.training_epochs = 40
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        losses = tr.training(_D())
        torch.cuda.synchronize()
        secs = time.perf_counter() - t0
        api = {"us_per_step": secs * 1e6 / (40 * 8), "steps": 40 * 8, "value": 40 * 8 / secs, "unit": "optimiser steps/s",
               "loss_first_epoch": losses[0], "loss_last_epoch": losses[-1],
               "api": "DynamicsTrainer.training(data set): wall clock of the whole call (moments, upload, 40 epochs x 8 batches of 512, "
                      "each epoch one CUDA-graph replay)"}
    flops = 3 * 2.0 * TRAIN_ROWS * 500 * 500 + 2 * 2.0 * TRAIN_ROWS * 500 * (6 + 2) * 1.5  # layer-2 fwd/dgrad/wgrad + the thin layers
    res = {"metric": "dynamics_train_steps_per_s", "value": steps / (ms * 1e-3), "unit": "optimiser steps/s",
           "samples_per_s": world * TRAIN_ROWS * steps / (ms * 1e-3), "us_per_step": ms * 1e3 / steps,
           "us_per_step_without_collective": ms_local * 1e3 / steps, "allreduce_us": ar_us,
           "rows_per_gpu": TRAIN_ROWS, "global_batch": gb, "scaling": "weak (global batch grows with N)",
           "network": "6->500->500->2 fp32 master weights; layer-2 GEMMs tcgen05 kind::f16 with fp16 hi/lo split operands (3 MMAs per product), fp32 accumulate",
           "bucket_floats": int(k.count + 1), "collective": "one NCCL all-reduce(sum) of the flat gradient bucket + loss per step" if world > 1 else None,
           "gpu_launches": launches, "launches_per_step": launches / steps, "loss_first": loss0, "loss_last": loss1,
           "tflops_per_gpu": flops / (ms * 1e-3 / steps) / 1e12}
    if p2p is not None:
        res["p2p"] = p2p
    if api is not None:
        res["training_api"] = api
    if with_cpu and rank == 0:
        # the reference's own path for this step: torch autograd + torch.optim.Adagrad on the host cores (oracle/train_oracle.py)
        from oracle.train_oracle import TorchBackend
        net_c = NNDynamics(4, 1, dxDim=2, oDim=5, xIsAngle=[False, True, False, False])
        cb = TorchBackend(net_c)
        g = lambda t: t.detach().cpu().numpy().ravel()
        cb.set_moments(g(net.oMean), g(net.oVar), g(net.uMean), g(net.uVar), g(net.dxMean), g(net.dxVar))
        cb.set_data(col)
        for i in range(3):
            cb.grad(batches[i][:TRAIN_ROWS], TRAIN_ROWS)
            cb.update(1e-3, 1e-3)
        n_cpu = 40
        t0 = time.perf_counter()
        for i in range(n_cpu):
            cb.grad(batches[i % 8][:TRAIN_ROWS], TRAIN_ROWS)
            cb.update(1e-3, 1e-3)
        secs = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": n_cpu / secs, "unit": "optimiser steps/s", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": "%d steps of batch %d through torch autograd + torch.optim.Adagrad on the host (the reference's own "
                                         "training path, mbrl.py:450-469)" % (n_cpu, TRAIN_ROWS)}
    return res


def main():
    ap = argparse.ArgumentParser()
    claim_stdout()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-ilqr", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-mlp", action="store_true")
    ap.add_argument("--no-nmpc", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from pygent_b200 import lib as pglib
    from pygent_b200.examples import make_ilqr
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the public objects a user would build: a batched CartPole and its iLQR (the compiled library is shared)
    x0_h, U_h = synthetic_inputs(seed=rank)
    alg = make_ilqr("cart_pole", ILQR_T, x0=np.zeros((2, 4)))   # only to obtain the problem library
    L = alg.lib
    x0 = x0_h.to(dev)
    U = U_h.to(dev)
    out = {"xN": torch.empty((4, B_ENVS), dtype=torch.float64, device=dev),
           "cost": torch.empty((B_ENVS,), dtype=torch.float64, device=dev),
           "terminated": torch.empty((B_ENVS,), dtype=torch.uint8, device=dev)}

    def step():
        L.rollout(x0, U, S=SUBSTEPS, dt=DT, out=out)

    # ---- device-resident throughput ------------------------------------------------------------------
    # clocks are sampled from before the warm-up until the last GPU phase ends: the timed region proper can be
    # shorter than one nvidia-smi period when the driver asks for few steps
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = L.launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for e0, e1 in evs:
        e0.record()
        step()
        e1.record()
    t_end.record()
    barrier()
    gpu_launches = L.launches - launches0
    elapsed_ms = t_start.elapsed_time(t_end)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    if world > 1:
        tmax = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tmax.item())
    value = world * B_ENVS * T_STEPS * args.steps / (elapsed_ms * 1e-3)

    # ---- the same workload in fp32 (FP32 FMA pipe view) ------------------------------------------------
    x0_32, U_32 = x0.float(), U.float()
    out32 = {"xN": torch.empty((4, B_ENVS), dtype=torch.float32, device=dev), "cost": torch.empty((B_ENVS,), dtype=torch.float32, device=dev),
             "terminated": torch.empty((B_ENVS,), dtype=torch.uint8, device=dev)}
    for _ in range(args.warmup):
        L.rollout(x0_32, U_32, S=SUBSTEPS, dt=DT, out=out32)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n32 = max(10, min(args.steps, 100))
    f0.record()
    for _ in range(n32):
        L.rollout(x0_32, U_32, S=SUBSTEPS, dt=DT, out=out32)
    f1.record()
    barrier()
    ms32 = f0.elapsed_time(f1) / n32
    err32 = float(((out32["xN"].double() - out["xN"]).abs() / (out["xN"].abs() + 1)).max())
    del x0_32, U_32

    # ---- SURVEY 8(d) variants of the resident workload: S = 1, and S = 4 with the full state history written ----------
    def time_variant(n, **kw):
        for _ in range(args.warmup):
            L.rollout(x0, U, dt=DT, **kw)
        barrier()
        v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        v0.record()
        for _ in range(n):
            L.rollout(x0, U, dt=DT, **kw)
        v1.record()
        barrier()
        ms = v0.elapsed_time(v1) / n
        if world > 1:
            tm = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ms = float(tm.item())
        return ms
    nv = max(10, min(args.steps, 50))
    ms_s1 = time_variant(nv, S=1, out=out)
    out_h = dict(out)
    out_h["X"] = torch.empty((T_STEPS + 1, 4, B_ENVS), dtype=torch.float64, device=dev)  # 1.05 GB
    ms_hist = time_variant(nv, S=SUBSTEPS, history=True, out=out_h)
    del out_h
    L.rollout(x0, U, S=SUBSTEPS, dt=DT, out=out)  # `out` holds the S = 4 results again (fp32 error, best-of-batch below)

    # ---- end to end through the public API, host buffers --------------------------------------------
    from pygent_b200.examples import make_env
    env = make_env("cart_pole", x0=np.ascontiguousarray(x0_h.numpy().T), device=dev)
    U_pin, x0_pin = U_h.pin_memory(), x0_h.pin_memory()
    cost_pin = torch.empty((B_ENVS,), dtype=torch.float64).pin_memory()
    xN_pin = torch.empty((4, B_ENVS), dtype=torch.float64).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step():
        r = env.rollout_host(x0_pin, U_pin, cost_out=cost_pin, xN_out=xN_pin)
        return r

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tmax = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s = float(tmax.item())
    e2e_host_value = world * B_ENVS * T_STEPS * e2e_steps / e2e_s
    e2e_host_gbs = (U_pin.numel() + x0_pin.numel()) * 8 * e2e_steps / e2e_s / 1e9   # per GPU
    assert torch.equal(xN_pin, out["xN"].cpu()), "e2e path and resident path disagree on the final states"
    assert torch.allclose(cost_pin, out["cost"].cpu(), rtol=1e-12, atol=0), "e2e path and resident path disagree"

    # ---- end to end with the actions drawn in the kernel ("or in-kernel Philox ... validated against the host
    # stream"): per step the host sends the start states (2 MB) and a seed, and reads cost and final states back ----------
    # Two sets of pinned buffers: step i + 1 is submitted before step i is waited for (`wait=False`), so the copies of one
    # step overlap the kernel of the next; every step still uploads its own start states and reads its own results back,
    # and the host touches each result (`best` below) before the buffer is reused.
    cost_pins = [torch.empty((B_ENVS,), dtype=torch.float64).pin_memory() for _ in range(2)]
    xN_pins = [torch.empty((4, B_ENVS), dtype=torch.float64).pin_memory() for _ in range(2)]
    x0_pins = [x0_pin, x0_pin.clone().pin_memory()]
    seed0 = 0x5EED0000 + rank
    for i in range(2):
        env.rollout_host_random(x0_pins[i], T_STEPS, seed0 + i, cost_out=cost_pins[i], xN_out=xN_pins[i])
    barrier()
    e2e_r_steps = max(3, min(args.steps, 50))
    best_seen = float("inf")
    t0 = time.perf_counter()
    pending = None
    for i in range(e2e_r_steps):
        h = env.rollout_host_random(x0_pins[i % 2], T_STEPS, seed0 + i, cost_out=cost_pins[i % 2], xN_out=xN_pins[i % 2], wait=False)
        if pending is not None:
            best_seen = min(best_seen, float(pending.wait()["cost"][0]))
        pending = h
    best_seen = min(best_seen, float(pending.wait()["cost"][0]))
    barrier()
    e2e_r_s = time.perf_counter() - t0
    cost_pin2, xN_pin2 = cost_pins[(e2e_r_steps - 1) % 2], xN_pins[(e2e_r_steps - 1) % 2]
    if world > 1:
        tmax = torch.tensor([e2e_r_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_r_s = float(tmax.item())
    e2e_value = world * B_ENVS * T_STEPS * e2e_r_steps / e2e_r_s
    # the stream is the documented one: the last step again over the materialised controls, and 512 environments of it
    # against the host restatement (pygent_b200/random_actions.py)
    from pygent_b200.random_actions import random_actions as host_actions
    U_chk = L.random_actions(B_ENVS, T_STEPS, seed0 + e2e_r_steps - 1, [U_MAX], device=dev)
    r_chk = L.rollout(x0, U_chk, S=SUBSTEPS, dt=DT)
    assert torch.equal(r_chk["xN"].cpu(), xN_pin2) and torch.equal(r_chk["cost"].cpu(), cost_pin2), "in-kernel action stream != materialised stream"
    assert np.array_equal(U_chk[:, :, :512].cpu().numpy(), host_actions(512, T_STEPS, seed0 + e2e_r_steps - 1, [U_MAX])), "device stream != host stream"
    del U_chk, r_chk

    # ---- final cost / argmin gather: the only collective of the path -----------------------
    from pygent_b200 import distributed as pgd
    bc, bi, br = pgd.gather_best(out["cost"], offset=rank * B_ENVS)
    best = {"cost": bc, "env": bi, "rank": br}

    # ---- iLQR iterations/s (config 3 sized per GPU) ---------------------------------------------------
    ilqr = None
    if not args.no_ilqr:
        rng = np.random.default_rng(1000 + rank)
        xi = np.array([0, np.pi, 0, 0]) + rng.uniform(-1, 1, (ILQR_B, 4)) * np.array([.5, .3, .5, .5])
        solver = make_ilqr("cart_pole", ILQR_T, x0=xi, env_kw={"device": dev}, maxIters=500)
        solver.run_optim()  # warm-up solve (same work)
        solver.environment.reset(xi)
        solver.reset()
        barrier()
        l0 = solver.lib.launches
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        solver.run_optim(host_results=False)   # results stay in HBM, like `value`; the copy to the host is timed below
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        total_iters = int(solver.iters.sum())
        solver.environment.reset(xi)
        solver.reset()
        torch.cuda.synchronize()
        a.record()
        solver.run_optim()                     # ... and with xx [B, T+1, n], uu [B, T, m] (66 MB) copied to host arrays
        b.record()
        torch.cuda.synchronize()
        ms_host = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms, -float(total_iters)], dtype=torch.float64, device=dev)
            tm = t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ts = t.clone()
            dist.all_reduce(ts, op=dist.ReduceOp.SUM)
            ms, total_iters = float(tm[0]), int(-ts[1])
        st = solver.status
        ilqr = {"metric": "ilqr_iterations_per_s", "value": total_iters / (ms * 1e-3), "unit": "iLQR iterations/s",
                "trajectories_per_gpu": ILQR_B, "horizon": ILQR_T, "substeps": SUBSTEPS, "ms": ms,
                "ms_with_host_results": ms_host,
                "iterations_total": total_iters, "iterations_max": int(solver.iters.max()),
                "gpu_launches": solver.lib.launches - l0,
                "converged_frac": float(np.mean((st == pglib.ST_CONV_FUN) | (st == pglib.ST_CONV_GRAD))),
                "mean_final_cost": float(np.mean(solver.cost)), "line_search": "serial semantics; first 4 of 11 alphas for all live trajectories, the other 7 only where none has passed yet "
                              "(one launch of all 11 once fewer than 64 x SMs trajectories are live)"}
        ilqr["full_batch"] = ilqr_full_batch_roofline(xi, dev)

    # ---- config 3 AS SPECIFIED: 16,384 initial states in total, sharded over the N GPUs (strong scaling) --------
    ilqr_c3 = None
    if not args.no_ilqr:
        rng = np.random.default_rng(1000)  # one global batch; rank r solves rows [r * B/N, (r + 1) * B/N)
        x_all = np.array([0, np.pi, 0, 0]) + rng.uniform(-1, 1, (ILQR_B, 4)) * np.array([.5, .3, .5, .5])
        per = ILQR_B // world
        xi = x_all[rank * per:(rank + 1) * per]
        solver = make_ilqr("cart_pole", ILQR_T, x0=xi, env_kw={"device": dev}, maxIters=500)
        solver.run_optim()
        solver.environment.reset(xi)
        solver.reset()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        solver.run_optim(host_results=False)
        b.record()
        barrier()
        ms, total_iters = a.elapsed_time(b), int(solver.iters.sum())
        if world > 1:
            t = torch.tensor([ms, -float(total_iters)], dtype=torch.float64, device=dev)
            tm, ts = t.clone(), t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(ts, op=dist.ReduceOp.SUM)
            ms, total_iters = float(tm[0]), int(-ts[1])
        bc3, bi3, br3 = pgd.gather_best(torch.as_tensor(solver.cost, device=dev), offset=rank * per)
        ilqr_c3 = {"metric": "ilqr_iterations_per_s", "value": total_iters / (ms * 1e-3), "unit": "iLQR iterations/s",
                   "trajectories_total": ILQR_B, "trajectories_per_gpu": per, "scaling": "strong", "horizon": ILQR_T, "ms": ms,
                   "iterations_total": total_iters, "best": {"cost": bc3, "trajectory": bi3, "rank": br3},
                   "note": "BASELINE.json configs[2]: 16,384 initial states in total, sharded over the GPUs of the run"}

    # ---- config 4: n = 6 double pendulum / acrobot, parallel line search -------------------------------
    ilqr_c4 = None
    if not args.no_ilqr:
        ilqr_c4 = {}
        for case, nominal in (("double_parallel", [0, np.pi, np.pi, 0, 0, 0]), ("acrobot", [np.pi, 0, 0, 0])):
            rng = np.random.default_rng(4000 + rank)
            xi = np.array(nominal) + rng.uniform(-0.1, 0.1, (8192, len(nominal)))
            solver = make_ilqr(case, ILQR_T, x0=xi, env_kw={"device": dev}, maxIters=100, parallel=True)
            solver.run_optim()
            solver.environment.reset(xi)
            solver.reset()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            solver.run_optim(host_results=False)
            b.record()
            barrier()
            ms, total_iters = a.elapsed_time(b), int(solver.iters.sum())
            if world > 1:
                t = torch.tensor([ms, -float(total_iters)], dtype=torch.float64, device=dev)
                tm, ts = t.clone(), t.clone()
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                dist.all_reduce(ts, op=dist.ReduceOp.SUM)
                ms, total_iters = float(tm[0]), int(-ts[1])
            ilqr_c4[case] = {"value": total_iters / (ms * 1e-3), "unit": "iLQR iterations/s", "trajectories_per_gpu": 8192,
                             "state_dim": len(nominal), "horizon": ILQR_T, "line_search": "parallel=True, 11 alphas", "ms": ms,
                             "iterations_total": total_iters, "mean_final_cost": float(np.mean(solver.cost))}

    # ---- receding-horizon control of a batch of plants ------------------------------
    nmpc = None
    if not args.no_nmpc:
        nmpc = nmpc_bench(dev, rank, world, barrier)

    # ---- iLQR through the learned-MLP dynamics (config 5 sized per GPU) --------------------------------
    mlp_ilqr = None
    if not args.no_mlp:
        mlp_ilqr = learned_ilqr_bench(dev, rank, world, barrier, precision="f16x3")   # the reference's precision: the config-5 result
        mlp_ilqr["lower_precision_bf16"] = learned_ilqr_bench(dev, rank, world, barrier, precision="bf16")

    # ---- dynamics training step + its gradient all-reduce --------------------------
    train = None
    if not args.no_train:
        train = train_bench(dev, rank, world, barrier, with_cpu=not args.no_cpu)

    clocks = sampler.stop()
    if rank == 0:
        peak_tf, peak_src = fp64_peak()
        hbm_gbs, hbm_src = hbm_peak()
        flops = algorithmic_flops_per_launch()
        achieved = flops / (kernel_ms * 1e-3) / 1e12
        hbm_bytes = U.numel() * 8 + x0.numel() * 8 + B_ENVS * (8 * 4 + 8 + 1)
        if ilqr and ilqr.get("full_batch"):
            fb = ilqr["full_batch"]
            fb["peak"], fb["frac"] = peak_tf, fb["achieved"] / peak_tf
        res = {
            "metric": "rk4_env_steps_per_s", "value": value, "unit": "env-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": CONFIG, "clocks": clocks, "gpu_launches": gpu_launches,
            "e2e": {"value": e2e_value, "unit": "env-steps/s",
                    "h2d_bytes_per_step": int(x0_pin.numel() * 8 + 8),
                    "d2h_bytes_per_step": int(cost_pin2.numel() * 8 + xN_pin2.numel() * 8), "steps": e2e_r_steps,
                    "api": "StateSpaceModel.rollout_host_random(x0_pinned, T, seed, wait=False): start states from pinned host memory, the random "
                           "actions of config 2 drawn inside the kernel (Philox2x32-10 stream, bit-identical to pygent_b200/random_actions.py "
                           "on the host: checked in this run), cost + final states back to pinned host memory; every step a new seed; "
                           "step i + 1 is submitted before step i's results are waited for (copies overlap the next kernel)"},
            "e2e_host_actions": {"value": e2e_host_value, "unit": "env-steps/s",
                                 "h2d_bytes_per_step": int(U_pin.numel() * 8 + x0_pin.numel() * 8),
                                 "d2h_bytes_per_step": int(cost_pin.numel() * 8 + xN_pin.numel() * 8), "steps": e2e_steps,
                                 "h2d_gb_per_s_per_gpu": e2e_host_gbs,
                                 "api": "StateSpaceModel.rollout_host(x0_pinned, U_pinned): the same workload with the 262 MB action array "
                                        "supplied by the host every step (PCIe-bound: compare h2d_gb_per_s_per_gpu with the link)"},
            "roofline": {"bound": "fp64", "kernel": "rollout_queue_kernel<Problem,double,RK4,1>", "achieved": achieved,
                         "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "peak_source": peak_src,
                         "flops_per_launch": flops, "kernel_ms": kernel_ms, "traffic": NCU_TRAFFIC_BYTES,
                         "note": "FP64 FMA pipe roofline (SURVEY 8d): algorithmic flops of the RK4 schedule (228 per substep) over "
                                 "the measured DFMA peak; the kernel issues fewer FP64 instructions than that count (sin/cos of "
                                 "the RK stages are rotations of one anchor per interval), pipe utilisation under ncu is 0.71",
                         "hbm": {"achieved": hbm_bytes / (kernel_ms * 1e-3) / 1e9, "peak": hbm_gbs, "unit": "GB/s",
                                 "frac": hbm_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_gbs, "bytes_per_launch": hbm_bytes,
                                 "peak_source": hbm_src}},
            "best": best,
            "s1": {"value": world * B_ENVS * T_STEPS / (ms_s1 * 1e-3), "unit": "env-steps/s", "substeps": 1, "ms_per_step": ms_s1,
                   "roofline": {"bound": "fp64", "achieved": B_ENVS * T_STEPS * (FLOPS_PER_SUBSTEP + FLOPS_COST) / (ms_s1 * 1e-3) / 1e12,
                                "peak": peak_tf, "unit": "TFLOP/s",
                                "frac": B_ENVS * T_STEPS * (FLOPS_PER_SUBSTEP + FLOPS_COST) / (ms_s1 * 1e-3) / 1e12 / peak_tf},
                   "hbm_gbs": hbm_bytes / (ms_s1 * 1e-3) / 1e9},
            "history": {"value": world * B_ENVS * T_STEPS / (ms_hist * 1e-3), "unit": "env-steps/s", "substeps": SUBSTEPS,
                        "ms_per_step": ms_hist, "bytes_per_env_step": 40,
                        "hbm_gbs": (hbm_bytes + 8.0 * 4 * B_ENVS * (T_STEPS + 1)) / (ms_hist * 1e-3) / 1e9,
                        "note": "X [T+1, n, B] fp64 (1.05 GB per GPU) written as well"},
            "fp32": {"value": world * B_ENVS * T_STEPS / (ms32 * 1e-3), "unit": "env-steps/s", "ms_per_step": ms32,
                     "roofline": {"bound": "fp32", "achieved": B_ENVS * T_STEPS * (SUBSTEPS * 188 + FLOPS_COST) / (ms32 * 1e-3) / 1e12,
                                  "peak": fp32_peak(), "unit": "TFLOP/s",
                                  "frac": B_ENVS * T_STEPS * (SUBSTEPS * 188 + FLOPS_COST) / (ms32 * 1e-3) / 1e12 / fp32_peak(),
                                  "flops_per_substep": 188},
                     "max_rel_state_error_vs_f64_after_500_intervals": err32,
                     "stated_tolerance": "5e-3 relative on states over 100 control intervals (tests/test_gpu_parity.py)"},
        }
        if ilqr is not None:
            res["ilqr"] = ilqr
        if mlp_ilqr is not None:
            res["mlp_ilqr"] = mlp_ilqr
        if nmpc is not None:
            res["nmpc"] = nmpc
        if ilqr_c4:
            res["ilqr_c4"] = ilqr_c4
        if ilqr_c3 is not None:
            res["ilqr_c3_sharded"] = ilqr_c3
        if train is not None:
            res["train"] = train
        if not args.no_cpu:
            import oracle
            oracle.build()
            threads = os.cpu_count() or 1
            n = B_ENVS  # the whole step: 65,536 envs x 500 control intervals (a few seconds on the host cores)
            v, secs = cpu_rollout_baseline(n, threads)
            v1, secs1 = cpu_rollout_baseline(1024, 1)
            if ilqr is not None:
                ilqr["cpu_baseline"] = cpu_ilqr_report(threads)
            res["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": threads, "kind": "port",
                                   "sample": "all %d envs x 500 steps of one step (RK4 x 4), C oracle, OpenMP over %d threads; %.1f s" % (n, threads, secs),
                                   "single_core": v1}
        emit(res)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
