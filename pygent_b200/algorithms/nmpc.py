"""Batched nonlinear MPC on the GPU behind pygent's `NMPC` / `MPCAgent` API (reference: pygent/algorithms/nmpc.py).

Every plant of the batch runs its own receding-horizon controller; one control step is, for all plants at once and
without a host synchronisation,

    step_iterations x (K3 backward, K2 forward over all alphas, K4 select, K2' commit)     the planner iteration
    pg_mpc_action      u = K_0 (x - xbar_0) + ubar_0 + alpha k_0 (+ noise), clipped         take_action  :187-211
    pg_mpc_shift       drop the first interval, append one with u = 0, fix the cost          shift_planner :257-277
    K1 pg_rollout      the controlled plant advances one interval                            NMPC.run :96-103

Reference behaviour that is kept on purpose: the gains are not shifted with the plan (only their last entry is
zeroed); the appended terminal state is integrated from wherever the planner's environment was left, i.e. the end of
the last line-search candidate evaluated, accepted or not; the transition cost of the shift uses t = None (0 here).
`finite_diff` follows the reference's defaults (True for a bare MPCAgent, nmpc.py:139; False through NMPC, nmpc.py:29).
Differences: `constrained=True` uses the exact-clip box QP
(DESIGN.md section 4), the noise generator is a seeded numpy RandomState owned by the agent.
"""
import os
import time

import numpy as np
import torch

from .. import lib
from ..agents import Agent
from ..helpers import OUnoise
from .core import Algorithm
from .ilqr import iLQR


class MPCAgent(Agent):
    def __init__(self, environment, horizon, dt, path,
                 init_iterations=500,
                 step_iterations=1,
                 fcost=None,
                 constrained=True,
                 fastForward=False,
                 tolGrad=1e-4,
                 tolFun=1e-7,
                 save_interval=10,
                 printing=True,
                 saving=True,
                 init_optim=True,
                 finite_diff=True,
                 noise_gain=0.05,
                 add_noise=False,
                 ou_theta=0.15,
                 ou_sigma=0.2,
                 seed=0,
                 cache_terminal_value=False):
        batch = environment.batch if environment.batched else None
        super(MPCAgent, self).__init__(environment.uDim, batch)
        self.traj_optimizer = iLQR(environment, horizon, dt,
                                   path=path if saving else None,
                                   fcost=fcost,
                                   constrained=constrained,
                                   fastForward=fastForward,
                                   save_interval=save_interval,
                                   tolGrad=tolGrad,
                                   tolFun=tolFun,
                                   printing=printing,
                                   finite_diff=finite_diff,
                                   file_prefix='ilqr_',
                                   init=False,
                                   saving=saving,
                                   reset_mu=True,
                                   cache_terminal_value=cache_terminal_value)
        self.uMax = np.asarray(self.traj_optimizer.environment.uMax, dtype=float)
        self.init_iterations = init_iterations
        self.step_iterations = step_iterations
        self.x0 = environment.x
        self.noise_gain = noise_gain
        self.add_noise = add_noise
        self.rng = np.random.RandomState(seed)
        self.action_noise = OUnoise(self.uDim, dt, theta=ou_theta, sigma=ou_sigma)
        self._u_dev = []       # applied controls, device tensors [m, B]
        self._xe = None        # state of the planner's environment (see pg_mpc_shift)
        if init_optim:
            self.init_optim(environment.x)

    # -- helpers -----------------------------------------------------------------------------------------------------
    def _x_dev(self, x):
        """state as a device tensor [n, B] (accepts one state, [B, n], or a tensor already in kernel layout)."""
        opt = self.traj_optimizer
        if torch.is_tensor(x) and x.is_cuda:
            return x
        return opt.environment._to_dev(np.asarray(x, dtype=float), opt.xDim)

    def _host_u(self, u):
        a = u.detach().to("cpu", torch.float64).numpy().T
        return a if self.batch is not None else a[0]

    def reset(self):
        super(MPCAgent, self).reset()
        self._u_dev = []

    @property
    def history(self):
        if self._u_dev:  # controls applied through the device path, zero dummy row in front (agents.py:23)
            uu = np.stack([self._host_u(u) for u in self._u_dev], axis=0 if self.batch is None else 1)
            zero = np.zeros_like(uu[:1] if self.batch is None else uu[:, :1])
            return np.concatenate([zero, uu], axis=0 if self.batch is None else 1)
        return super(MPCAgent, self).history

    # -- the reference's methods -------------------------------------------------------------------------------------
    def init_optim(self, x0, init=True):
        """Full optimisation from x0 (nmpc.py:174-184)."""
        opt = self.traj_optimizer
        opt.reset_mu = True
        opt.environment.reset(x0 if not torch.is_tensor(x0) else x0.detach().cpu().numpy().T)
        opt.environment.x0 = opt.environment.x
        opt.init = init
        opt.reset()
        opt.max_iters = self.init_iterations
        opt.run_optim()
        opt.max_iters = self.step_iterations
        opt.reset_mu = True
        # the planner's environment is left at the end of the last candidate evaluated
        lt = opt._last_tried.long().clamp(min=0)
        idx = lt[None, None, :].expand(1, opt.xDim, opt.B)
        cand = opt._cand["X"][:, opt.steps] if opt.learned else opt._fw["xN"]
        self._xe = torch.gather(cand, 0, idx)[0].contiguous()

    def load_plan(self, xx, uu, KK, kk, cost, alpha, env_x=None):
        """Warm start from a stored plan / controller (arrays shaped like `iLQR.xx, uu, KK, kk`, one plant or a batch):
        what `init_optim` leaves behind, without running it.  env_x: state of the planner's environment (defaults to the
        plan's terminal state)."""
        opt = self.traj_optimizer
        B, T, n, m = opt.B, opt.steps, opt.xDim, opt.uDim
        dev = lambda a, shape: torch.as_tensor(np.ascontiguousarray(np.moveaxis(np.asarray(a, dtype=float).reshape((B,) + shape), 0, -1)),
                                               dtype=opt.dtype, device=opt.environment.device)
        opt._xbar.copy_(dev(xx, (T + 1, n)))
        opt._ubar.copy_(dev(uu, (T, m)))
        opt._KK.copy_(dev(KK, (T, m, n)))
        opt._kk.copy_(dev(kk, (T, m)))
        opt._st["cost"].copy_(dev(cost, ()))
        opt._st["alpha_best"].copy_(dev(alpha, ()))
        opt._have_gains = True
        opt._last_tried.fill_(-1)
        self._xe = dev(env_x if env_x is not None else np.asarray(xx, dtype=float).reshape(B, T + 1, n)[:, -1], (n,))
        opt.max_iters = self.step_iterations
        opt.reset_mu = True

    def noise(self):
        shape = (self.uDim,) if self.batch is None else (self.batch, self.uDim)
        return self.uMax * self.rng.normal(0, self.noise_gain, shape)

    def take_action_device(self, x):
        """One control step for every plant, device only.  x: measured states [n, B]; returns u [m, B]."""
        opt = self.traj_optimizer
        opt._x0 = x
        opt.max_iters = self.step_iterations
        opt.run_iterations(self.step_iterations)
        noise = None
        if self.add_noise:
            noise = opt.environment._to_dev(self.noise(), self.uDim)
        u = opt.lib.mpc_action(opt._KK, opt._kk, opt._xbar, opt._ubar, opt._st["alpha_best"], x, self.uMax, noise=noise)
        self._u_dev.append(u)
        self.tt.extend([self.tt[-1] + opt.dt])
        self.shift_planner()
        return u

    def take_action(self, dt, x):
        """Control for the measured state(s) x (nmpc.py:187-211); returns a host array like the reference."""
        self.u = self._host_u(self.take_action_device(self._x_dev(x)))
        return self.u

    def _plan_host(self):
        """host copy of the plan (one device read per plan, not per control step)"""
        opt = self.traj_optimizer
        key = (opt.lib.launches, id(opt))
        if getattr(self, "_plan_cache", (None,))[0] != key:
            self._plan_cache = (key, opt.xx, opt.uu, opt.KK, opt.kk, np.asarray(opt.current_alpha, dtype=float))
        return self._plan_cache[1:]

    def _plan_action(self, i, x=None):
        xx, uu, KK, kk, alpha = self._plan_host()
        u = uu[..., i, :] + np.reshape(alpha, (-1, 1) if self.batch is not None else ()) * kk[..., i, :, 0]
        if x is not None:
            K = KK[..., i, :, :]
            u = u + np.einsum("...mn,...n->...m", K, np.asarray(x, dtype=float) - xx[..., i, :])
        u = u + self.add_noise * self.noise()
        self.u = np.clip(u, -self.uMax, self.uMax)
        self.control(self.traj_optimizer.dt, self.u)
        return self.u

    def take_action_plan(self, dt, i):
        """Open-loop plan + feed-forward term (nmpc.py:214-231)."""
        return self._plan_action(i)

    def take_action_plan_feedback(self, dt, x, i):
        """Plan with state feedback (nmpc.py:233-254)."""
        i = min(self.traj_optimizer.steps - 1, i)
        return self._plan_action(i, x)

    def shift_planner(self):
        """nmpc.py:257-277 on the device (pg_mpc_shift)."""
        opt = self.traj_optimizer
        env = opt.environment
        t_env = opt.steps * opt.dt
        if opt.learned:
            lt = opt._last_tried.long()
            idx = lt.clamp(min=0)[None, None, :].expand(1, opt.xDim, opt.B)
            xs = torch.where(lt[None, :] >= 0, torch.gather(opt._cand["X"][:, opt.steps], 0, idx)[0], self._xe).contiguous()
            x_new = env.dynamics.rollout_device(xs, opt.dt, T=1, precision=env.precision)[1].contiguous()
            self._xe.copy_(xs)
            opt.lib.mpc_shift(opt._xbar, opt._ubar, opt._KK, opt._kk, opt._st["cost"], self._xe, dt=opt.dt, S=1,
                              integrator=lib.INTEG_EULER, t_env=t_env, x_new_ext=x_new,
                              terminal_cost=float(env.terminal_cost))
        else:
            # the reference steps its planner environment with `step` (the RK integrator) even when fastForward is set
            opt.lib.mpc_shift(opt._xbar, opt._ubar, opt._KK, opt._kk, opt._st["cost"], self._xe, dt=opt.dt, S=opt.substeps,
                              integrator=lib.INTEG_RK4, t_env=t_env, xN_cand=opt._fw["xN"], last_tried=opt._last_tried,
                              terminal_cost=float(env.terminal_cost))

    def take_random_action(self, dt, ou_noise=True):
        """nmpc.py:279-296."""
        if ou_noise:
            u = self.action_noise.sample() * self.uMax
            self.u = np.clip(u, -self.uMax, self.uMax)
        else:
            shape = (self.uDim,) if self.batch is None else (self.batch, self.uDim)
            self.u = self.rng.uniform(-self.uMax, self.uMax, shape)
        self.control(dt, self.u)
        return self.u


class NMPC(Algorithm):
    """Closed loop of `environment` (the plant) with an `MPCAgent` planning on `mpc_environment` (nmpc.py:16-103)."""

    def __init__(self, environment, mpc_environment, t, dt, horizon,
                 maxIters=500,
                 step_iterations=1,
                 tolGrad=1e-4,
                 tolFun=1e-7,
                 fastForward=False,
                 path=None,
                 fcost=None,
                 constrained=True,
                 save_interval=50,
                 init_optim=True,
                 finite_diff=False,
                 ilqr_print=False,
                 ilqr_save=False,
                 noise_gain=0.005,
                 add_noise=False,
                 ou_theta=0.15,
                 ou_sigma=0.2,
                 cache_terminal_value=False):
        self.path = path
        if path is not None:
            for sub in ("", "plots/", "animations/", "data/"):
                os.makedirs(path + sub, exist_ok=True)
        self.mpc_environment = mpc_environment
        agent = MPCAgent(mpc_environment, horizon, dt, path,
                         init_iterations=maxIters,
                         step_iterations=step_iterations,
                         tolGrad=tolGrad,
                         tolFun=tolFun,
                         fcost=fcost,
                         constrained=constrained,
                         fastForward=fastForward,
                         printing=ilqr_print,
                         saving=ilqr_save and path is not None,
                         init_optim=init_optim,
                         finite_diff=finite_diff,
                         noise_gain=noise_gain,
                         add_noise=add_noise,
                         save_interval=save_interval,
                         ou_sigma=ou_sigma,
                         ou_theta=ou_theta,
                         cache_terminal_value=cache_terminal_value)
        super(NMPC, self).__init__(environment, agent, t, dt)
        self.save_interval = save_interval

    def _graph_ok(self):
        """One control step can be replayed as a CUDA graph when nothing in it needs the host: no action noise (numpy
        generator), the planner's terminal value cached (or no final backward pass), and costs that do not depend on time
        (the plant's clock is a launch argument)."""
        import sympy as sp
        opt = self.agent.traj_optimizer
        t = sp.Symbol("t")
        time_free = t not in sp.sympify(self.environment.cost_expression()).free_symbols and \
            t not in sp.sympify(opt.environment.cost_expression()).free_symbols
        return (not self.agent.add_noise) and time_free and (opt.cache_terminal_value or not opt.final_backpass)

    def run(self, steps=None, printing=True, use_graph=None):
        """Closed-loop simulation (nmpc.py:87-103): the loop runs on the device; plant states, controls and costs come
        back once at the end.  Returns the per-step transition costs `[steps]` / `[B, steps]`.
        use_graph (default: when possible, see `_graph_ok`): after two eager control steps the whole step -- planner
        iteration, final backward pass, action, plan shift, plant step, state hand-over: ~25 launches -- is captured into
        a CUDA graph and replayed."""
        start_time = time.time()
        env, agent = self.environment, self.agent
        env.reset(env.x0)
        agent.reset()
        n_steps = len(np.arange(0, self.t, self.dt)) if steps is None else int(steps)
        L = env.library()
        x = env._state()
        B = env.batch
        X = torch.empty((n_steps + 1, env.xDim, B), dtype=env.dtype, device=env.device)
        C = torch.empty((n_steps, B), dtype=env.dtype, device=env.device)
        term = torch.empty((B,), dtype=torch.uint8, device=env.device)
        X[0].copy_(x)
        t_plant = 0.0
        if use_graph is None:
            # pays where the step is launch-bound on the host: measured at 4,096 plants the ~25 kernels of a step take
            # 0.45 ms of GPU time and eager launching keeps up (0.49 ms per step), the capture itself costs ~15 ms
            # replaying costs the GPU nothing and takes the host out of the loop: at 4,096 plants the eager loop is
            # GPU-bound on a quiet host (0.44 ms per step) but 0.57-1.09 ms when the host is busy; the replay stays at 0.45
            use_graph = self._graph_ok() and n_steps >= 50 and os.environ.get("PG_NMPC_GRAPH", "1") != "0"
        elif use_graph and not self._graph_ok():
            raise ValueError("this controller needs the host inside a control step (noise, time-dependent cost, or an uncached "
                             "final backward pass): it cannot be captured into a CUDA graph")
        eager = min(n_steps, 2) if use_graph else n_steps
        for i in range(eager):
            u = agent.take_action_device(X[i])
            L.rollout(X[i], u[None], S=env.substeps, dt=self.dt, t0=t_plant, terminal_cost=float(env.terminal_cost),
                      out={"xN": X[i + 1], "cost": C[i], "terminated": term})
            t_plant = t_plant + self.dt
        if eager < n_steps:
            K = max(1, min(int(os.environ.get("PG_NMPC_GRAPH_STEPS", "10")), n_steps - eager))
            libs = list({id(l): l for l in (L, agent.traj_optimizer.lib)}.values())
            g = getattr(self, "_graph_state", None)
            if g is None or g["K"] != K or g["B"] != B or g["planner"] is not agent.traj_optimizer:
                # K consecutive control steps per graph: one replay + three copies per K steps on the host instead of per
                # step, so a busy host (8 ranks, clock samplers) does not pace the loop.  Captured on a side stream with
                # the raw API: `with torch.cuda.graph(...)` would first empty the caching allocator (hundreds of ms of
                # cudaFree when another solver has just released its buffers).  The graph is kept for later run() calls.
                x_in = X[eager].clone()
                Xs = torch.empty((K, env.xDim, B), dtype=env.dtype, device=env.device)
                Cs = torch.empty((K, B), dtype=env.dtype, device=env.device)
                Us, gterm = None, torch.empty((B,), dtype=torch.uint8, device=env.device)
                n_before, tt_before = len(agent._u_dev), len(agent.tt)
                launched = [l.launches for l in libs]
                graph = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream(device=env.device)
                side.wait_stream(torch.cuda.current_stream(env.device))
                with torch.cuda.stream(side), lib.workspace_stream(torch.cuda.current_stream(env.device)):
                    graph.capture_begin(capture_error_mode="thread_local")  # other threads (NCCL watchdog, samplers) may call CUDA
                    try:
                        for k in range(K):
                            u_g = agent.take_action_device(x_in)
                            if Us is None:
                                Us = torch.empty((K,) + tuple(u_g.shape), dtype=u_g.dtype, device=u_g.device)
                            L.rollout(x_in, u_g[None], S=env.substeps, dt=self.dt, t0=0.0, terminal_cost=float(env.terminal_cost),
                                      out={"xN": Xs[k], "cost": Cs[k], "terminated": gterm})
                            Us[k].copy_(u_g)
                            x_in.copy_(Xs[k])
                    finally:
                        graph.capture_end()
                torch.cuda.current_stream(env.device).wait_stream(side)
                del agent._u_dev[n_before:]  # the capture recorded (did not run) K steps: undo their host-side bookkeeping
                del agent.tt[tt_before:]
                per_replay = [l.launches - n0 for l, n0 in zip(libs, launched)]
                for l, n0 in zip(libs, launched):
                    l.launches = n0
                g = self._graph_state = {"K": K, "B": B, "planner": agent.traj_optimizer, "graph": graph, "x_in": x_in, "Xs": Xs, "Cs": Cs,
                                         "Us": Us, "term": gterm, "per_replay": per_replay}
            else:
                g["x_in"].copy_(X[eager])
            graph, Xs, Cs, Us = g["graph"], g["Xs"], g["Cs"], g["Us"]
            n_rep = (n_steps - eager) // K
            Uh = torch.empty((n_rep * K,) + tuple(Us.shape[1:]), dtype=Us.dtype, device=Us.device)
            i = eager
            for r in range(n_rep):
                graph.replay()
                X[i + 1:i + 1 + K].copy_(Xs)
                C[i:i + K].copy_(Cs)
                Uh[r * K:(r + 1) * K].copy_(Us)
                for k in range(K):
                    agent._u_dev.append(Uh[r * K + k])
                    agent.tt.extend([agent.tt[-1] + self.dt])
                i += K
            for l, per in zip(libs, g["per_replay"]):
                l.launches += per * n_rep
            if n_rep:
                term.copy_(g["term"])
                t_plant = t_plant + self.dt * n_rep * K
            for i in range(i, n_steps):  # the remainder that does not fill a graph: eager control steps
                u = agent.take_action_device(X[i])
                L.rollout(X[i], u[None], S=env.substeps, dt=self.dt, t0=t_plant, terminal_cost=float(env.terminal_cost),
                          out={"xN": X[i + 1], "cost": C[i], "terminated": term})
                t_plant = t_plant + self.dt
        Xh = X.detach().to("cpu", torch.float64).numpy()
        env._hist = [Xh[k].T.copy() for k in range(n_steps + 1)]
        env._xd_prev, env._xd = X[n_steps - 1].clone(), X[n_steps].clone()
        for _ in range(n_steps):
            env.tt.append(env.tt[-1] + self.dt)
        tm = term.cpu().numpy().astype(bool)
        env.terminated = tm if env.batched else bool(tm[0])
        cost = C.detach().to("cpu", torch.float64).numpy().T
        cost = cost if env.batched else cost[0]
        if printing:
            print('Cost: %.2f | Runtime: %.2f min.' % (float(np.mean(np.sum(cost, axis=-1))), (time.time() - start_time) / 60))
        if self.path is not None:
            env.save_history('environment', self.path + 'data/')
            agent.save_history('agent', self.path + 'data/')
        return cost
