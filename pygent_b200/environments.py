"""Batched environments behind pygent's `Environment` API, integrated on the GPU.

Mirrors `pygent/environments.py` (reference): `Environment` (:23-120), `StateSpaceModel` (:182-326)
and the concrete systems `Pendulum` (:328-357), `AngularPendulum` (:394-423), `CartPole` (:461-482),
`CartPoleDoubleSerial` (:523-551), `CartPoleDoubleParallel` (:596-615), `CartPoleTriple` (:660-679),
`Car` (:730-761), `Acrobot` (:763-782), `MarBot` (:826-868) -- same constructor signatures, attribute
names (`x, x_, x0, xDim, uDim, oDim, o, o_, xIsAngle, uMax, dt, history, tt, terminated`) and methods
(`reset, step, fast_step, terminate, observe, ode, A, B`).

What is different by design:
  * `x0` may be one state (list of n) or a batch `[B, n]`; all B environments advance in one kernel
    launch.  With a single state the attributes have the reference's shapes (`x` is `[n]`,
    `history` is `[k+1, n]`); with a batch they gain a leading B axis.
  * `step` integrates with fixed-step RK4, `substeps` per control interval (default 4: within 3e-9 of
    the reference's adaptive RK45 on these systems) -- the parity schedule.
  * `rollout(U)` runs a whole control sequence in ONE launch (the reference has only Python loops
    over `step`, e.g. `pygent/algorithms/ilqr.py:504-512`).
State lives on the GPU in trajectory-minor layout `[n, B]`; there is no CPU integration path.
"""
import numpy as np
import sympy as sp
import torch

from . import codegen, lib, tracing
from .helpers import observation
from .modeling.systems import SystemModel, get_system


def _default_device():
    if not torch.cuda.is_available():
        raise lib.PygentB200Error("pygent_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class Environment(object):
    """Environment base class (reference: environments.py:23-120)."""

    def __init__(self, x0, uDim, dt, device=None, dtype=torch.float64):
        if callable(x0):
            self.x0 = x0
            x0 = x0()
        else:
            x0 = np.array(x0, dtype=float)
            self.x0 = x0
        x0 = np.array(x0, dtype=float)
        self.batched = x0.ndim == 2
        self.batch = x0.shape[0] if self.batched else 1
        self.xDim = x0.shape[-1]
        self.oDim = self.xDim
        self.uDim = uDim
        self.xIsAngle = np.zeros([self.xDim], dtype=bool)
        self.uMax = np.ones(uDim)
        self.dt = dt
        self.dtype = dtype
        self._device = device
        self._set_state(x0)

    # -- device state ---------------------------------------------------------------------------
    @property
    def device(self):
        if self._device is None:
            self._device = _default_device()
        return self._device

    def _to_dev(self, a, cols):
        """host [cols] or [B, cols] -> device [cols, B]"""
        a = np.asarray(a, dtype=float).reshape(self.batch, cols)
        return torch.as_tensor(a.T.copy(), dtype=self.dtype, device=self.device).contiguous()

    def _to_host(self, t):
        """device [c, B] -> numpy [c] (single) or [B, c]"""
        a = t.detach().to("cpu", torch.float64).numpy().T
        return a if self.batched else a[0]

    def _set_state(self, x0):
        self._x0_host = np.array(x0, dtype=float).reshape(self.batch, self.xDim)
        self._xd = None          # device state, created lazily (so that construction works without a GPU)
        self._xd_prev = None
        self._hist = [self._x0_host.copy()]
        if "_obs_cur_state" not in self.__dict__:   # `o` is computed at construction and by `step`, never by `reset` (see `o_`)
            self._obs_cur_state = self._hist[0]
        self.tt = [0]
        self.terminated = False if not self.batched else np.zeros(self.batch, dtype=bool)

    def _state(self):
        if self._xd is None:
            self._xd = self._to_dev(self._x0_host, self.xDim)
            self._xd_prev = self._xd
        return self._xd

    @property
    def x(self):
        return self._to_host(self._state()) if self._xd is not None else (self._x0_host if self.batched else self._x0_host[0])

    @property
    def x_(self):
        return self._to_host(self._xd_prev) if self._xd is not None else self.x

    @property
    def history(self):
        h = np.stack(self._hist, axis=1)  # [B, k+1, n]
        return h if self.batched else h[0]

    def get_state(self):
        return self.x

    def reset(self, x0):
        """Resets environment to state x0 (list of n, `[B, n]` array, or callable)."""
        if callable(x0):
            x0 = x0()
        x0 = np.array(x0, dtype=float)
        if x0.ndim == 2 and x0.shape[0] != self.batch:
            self.batch, self.batched = x0.shape[0], True
        self._set_state(x0)

    def step(self, *args):
        raise NotImplementedError

    def observe(self, x):
        return x

    # plotting / animation: out of scope; accepted, do nothing (environments.py:81-126)
    def plot(self, *args, **kwargs):
        return None

    def animation(self, *args, **kwargs):
        return None

    def save_history(self, filename, path):
        import pickle
        with open(path + filename + ".p", "wb") as fh:
            pickle.dump({"tt": self.tt, "xx": self.history}, fh)


class StateSpaceModel(Environment):
    """dx/dt = f(x, u) environment (reference: environments.py:182-326).

    `ode` is a `SystemModel`, a registered system name, or a callable `ode(t, x, u)` that is traced
    into sympy (numpy ufuncs inside it are fine)."""

    def __init__(self, ode, cost, x0, uDim, dt, terminal_cost=0., substeps=4, device=None, dtype=torch.float64,
                 name=None):
        super(StateSpaceModel, self).__init__(x0, uDim, dt, device=device, dtype=dtype)
        if isinstance(ode, str):
            ode = get_system(ode)
        if isinstance(ode, SystemModel):
            self.system = ode
        else:
            xs = sp.symbols("x1:%d" % (self.xDim + 1))
            us = sp.symbols("u1:%d" % (uDim + 1))
            f = tracing.trace(ode, None, list(xs), list(us))
            self.system = SystemModel(name or "user_%s" % getattr(ode, "__name__", "ode"), f, xs, us, False)
        if self.system.n != self.xDim or self.system.m != uDim:
            raise ValueError("system %s has n=%d, m=%d but x0/uDim give n=%d, m=%d"
                             % (self.system.name, self.system.n, self.system.m, self.xDim, uDim))
        self._user_cost = cost
        self.cost = tracing.bind_cost(cost)
        self.terminal_cost = terminal_cost
        self.substeps = substeps
        self.xIsAngle = np.zeros([self.xDim], dtype=bool)
        self.oDim = self.xDim
        self._problems = {}
        self._cost_expr = None

    def __getattr__(self, name):
        # numpy callables `ode(t, x, u)`, `A(x, u)`, `B(x, u)` of the reference API, lambdified on first use;
        # like the reference, only systems with symbolic Jacobians expose A and B (ilqr.py:523)
        if name in ("ode", "A", "B") and "system" in self.__dict__:
            f_np, A_np, B_np = self.system.numpy_callables()
            if name == "ode":
                return f_np
            if self.system.has_AB:
                return A_np if name == "A" else B_np
        raise AttributeError(name)

    # -- problem compilation ----------------------------------------------------------------------
    def _term_limits(self):
        """[(state index, limit)]: terminated = any(|x_i| > limit); subclasses fill this in."""
        return []

    def cost_expression(self):
        """Running cost as a sympy expression in (x, u, x_next, t), obtained like `iLQR.cost_init`
        does (ilqr.py:539-543): the user's callable evaluated on symbols with mod = sympy."""
        if self._cost_expr is None:
            s = self.system
            xn = sp.symbols("xn1:%d" % (s.n + 1))
            self._cost_expr = tracing.trace(self.cost, list(s.xs), list(s.us), list(xn), sp.Symbol("t"), sp)
        return self._cost_expr

    def problem_spec(self, fcost_expr=None):
        return codegen.ProblemSpec(self.system, self.cost_expression(), fcost_expr, term=self._term_limits(),
                                   x_is_angle=list(self.xIsAngle))

    def library(self, fcost_expr=None):
        """The compiled CUDA library of (dynamics, running cost, final cost); built on first use."""
        key = sp.srepr(sp.sympify(fcost_expr)) if fcost_expr is not None else None
        key = (key, tuple(self._term_limits()), tuple(bool(a) for a in self.xIsAngle))
        if key not in self._problems:
            self._problems[key] = lib.load_problem(self.problem_spec(fcost_expr))
        return self._problems[key]

    # -- stepping ---------------------------------------------------------------------------------
    def _advance(self, args, integrator):
        if len(args) == 2:
            u, dt = args
        elif len(args) == 1:
            u, dt = args[0], self.dt
        else:
            raise TypeError("step(u[, dt])")
        L = self.library()
        x = self._state()
        ud = self._to_dev(u, self.uDim).reshape(1, self.uDim, self.batch)
        r = L.rollout(x, ud, S=self.substeps, dt=dt, t0=float(self.tt[-1]), integrator=integrator,
                      terminal_cost=float(self.terminal_cost))
        self._xd_prev, self._xd = x, r["xN"]
        self._hist.append(r["xN"].detach().to("cpu", torch.float64).numpy().T.copy())
        # `o_ = o; ...; o = observe(x)` (environments.py:254-267, :311-316): the state whose observation `o_` reports is the one
        # `o` was last computed from -- the previous step's result, or x0 at construction -- NOT necessarily x_ (see `o_`)
        self._obs_prev_state = self._obs_cur_state
        self._obs_cur_state = self._hist[-1]
        self.tt.extend([self.tt[-1] + dt])
        term = r["terminated"].cpu().numpy().astype(bool)
        c = r["cost"].to("cpu", torch.float64).numpy()
        if self.batched:
            self.terminated = term
            return c
        self.terminated = bool(term[0])
        return float(c[0])

    def step(self, *args):
        """One control interval `dt` with control `u` (`[m]` or `[B, m]`); returns the transition cost
        `(cost(x_, u, x, t) + terminal_cost*terminated)*dt` (reference: environments.py:242-276, with
        solve_ivp replaced by RK4 x `substeps`)."""
        return self._advance(args, lib.INTEG_RK4)

    def fast_step(self, *args):
        """Forward-Euler step (reference: environments.py:291-322)."""
        return self._advance(args, lib.INTEG_EULER)

    def rollout(self, U, history=True, fast=False, add_final_cost=False, fcost_expr=None):
        """Whole control sequence in one launch.  U: `[T, m]` / `[B, T, m]` host array, or a device
        tensor already in kernel layout `[T, m, B]`.  Equivalent to `for u in U: step(u)`.
        Returns dict(X, xN, cost, terminated) as device tensors in kernel layout."""
        L = self.library(fcost_expr)
        x = self._state()
        if torch.is_tensor(U) and U.is_cuda:
            Ud = U
        else:
            U = np.asarray(U, dtype=float)
            if U.ndim == 2:
                U = np.broadcast_to(U[None], (self.batch,) + U.shape)
            Ud = torch.as_tensor(np.array(U.transpose(1, 2, 0), order="C"), dtype=self.dtype, device=self.device)
        T = Ud.shape[0]
        r = L.rollout(x, Ud, S=self.substeps, dt=self.dt, t0=float(self.tt[-1]),
                      integrator=lib.INTEG_EULER if fast else lib.INTEG_RK4, history=history,
                      terminal_cost=float(self.terminal_cost), add_final_cost=add_final_cost)
        self._xd_prev, self._xd = x, r["xN"]
        if history:
            Xh = r["X"].detach().to("cpu", torch.float64).numpy()  # [T+1, n, B]
            self._hist.extend([Xh[k].T.copy() for k in range(1, T + 1)])
        else:
            self._hist.append(r["xN"].detach().to("cpu", torch.float64).numpy().T.copy())
        t = self.tt[-1]
        for _ in range(T):
            t = t + self.dt
            self.tt.append(t)
        term = r["terminated"].cpu().numpy().astype(bool)
        L.check_queues()  # we have just synchronised: a queue worker that gave up must not go unnoticed
        self.terminated = term if self.batched else bool(term[0])
        return r

    def rollout_host(self, x0, U, cost_out=None, xN_out=None, segments=10, fast=False):
        """End-to-end rollout from HOST buffers in kernel layout: x0 `[n, B]`, U `[T, m, B]` (pinned torch
        tensors), results into the pinned `cost_out [B]` / `xN_out [n, B]`.  The horizon is fed in `segments`
        time slices: while slice k integrates, slice k+1 crosses PCIe on a second stream (double buffered),
        so the wall time approaches max(copy, compute) instead of their sum.  Synchronises before returning."""
        L = self.library()
        T, m, B = U.shape
        n = self.xDim
        dev, dt_ = self.device, self.dtype
        seg = (T + segments - 1) // segments
        st = getattr(self, "_host_pipe", None)
        if st is None or st["shape"] != (seg, m, B):
            st = {"shape": (seg, m, B), "copy": torch.cuda.Stream(device=dev),
                  "buf": [torch.empty((seg, m, B), dtype=dt_, device=dev) for _ in range(2)],
                  "x": [torch.empty((n, B), dtype=dt_, device=dev) for _ in range(2)],
                  "cost": torch.empty((B,), dtype=dt_, device=dev),
                  "term": torch.empty((B,), dtype=torch.uint8, device=dev)}
            self._host_pipe = st
        comp = torch.cuda.current_stream(dev)
        copy = st["copy"]
        copy.wait_stream(comp)
        with torch.cuda.stream(copy):
            st["x"][0].copy_(x0, non_blocking=True)
        done = [None, None]   # compute-done events per U buffer
        t0, k, i = 0.0, 0, 0
        while k < T:
            k1 = min(T, k + seg)
            b = i % 2
            with torch.cuda.stream(copy):
                if done[b] is not None:
                    copy.wait_event(done[b])
                st["buf"][b][: k1 - k].copy_(U[k:k1], non_blocking=True)
                ready = copy.record_event()
            comp.wait_event(ready)
            L.rollout(st["x"][i % 2], st["buf"][b][: k1 - k], S=self.substeps, dt=self.dt, t0=t0,
                      integrator=lib.INTEG_EULER if fast else lib.INTEG_RK4, terminal_cost=float(self.terminal_cost),
                      accumulate_cost=i > 0, out={"xN": st["x"][(i + 1) % 2], "cost": st["cost"], "terminated": st["term"]})
            done[b] = comp.record_event()
            for _ in range(k1 - k):
                t0 = t0 + self.dt
            k, i = k1, i + 1
        if cost_out is not None:
            cost_out.copy_(st["cost"], non_blocking=True)
        if xN_out is not None:
            xN_out.copy_(st["x"][i % 2], non_blocking=True)
        comp.synchronize()
        L.check_queues()
        return {"cost": cost_out, "xN": xN_out}

    def rollout_host_random(self, x0, T, seed, cost_out=None, xN_out=None, u_max=None, fast=False, wait=True):
        """End-to-end rollout under RANDOM actions from HOST start states: x0 `[n, B]` (pinned torch tensor) crosses PCIe,
        the T controls per environment are drawn inside the kernel (Philox stream of pygent_b200/random_actions.py, keyed
        by `seed`: u ~ U(-uMax, uMax)) -- 2 MB of input per step at the benchmark size instead of the
        264 MB a materialised U would be -- and cost `[B]` / final states `[n, B]` come back into the pinned outputs.
        `random_actions.random_actions(B, T, seed, uMax)` reproduces the controls on the host.

        `wait=True` synchronises before returning.  `wait=False` returns a handle whose `.wait()` blocks until THIS call's
        outputs are in host memory: copies in, kernel and copies out run on three streams over two alternating sets of
        device buffers, so a caller that submits step i + 1 before waiting for step i overlaps the 4.7 MB of copies of one
        step with the kernel of the next (give consecutive calls different output buffers)."""
        L = self.library()
        n, B = x0.shape
        dev, dt_ = self.device, self.dtype
        st = getattr(self, "_host_rand", None)
        if st is None or st["B"] != B:
            slot = lambda: {"x": torch.empty((n, B), dtype=dt_, device=dev), "xN": torch.empty((n, B), dtype=dt_, device=dev),
                            "cost": torch.empty((B,), dtype=dt_, device=dev), "term": torch.empty((B,), dtype=torch.uint8, device=dev),
                            "free": None}
            st = self._host_rand = {"B": B, "slots": [slot(), slot()], "next": 0, "h2d": torch.cuda.Stream(device=dev),
                                    "d2h": torch.cuda.Stream(device=dev)}
        um = np.ravel(self.uMax if u_max is None else u_max).astype(float)
        sl = st["slots"][st["next"]]
        st["next"] ^= 1
        comp = torch.cuda.current_stream(dev)
        with torch.cuda.stream(st["h2d"]):
            if sl["free"] is not None:
                st["h2d"].wait_event(sl["free"])     # the slot's previous results have left the device
            sl["x"].copy_(x0, non_blocking=True)
            up = torch.cuda.Event()
            up.record(st["h2d"])
        comp.wait_event(up)
        L.rollout(sl["x"], None, T=int(T), S=self.substeps, dt=self.dt, t0=0.0, integrator=lib.INTEG_EULER if fast else lib.INTEG_RK4,
                  terminal_cost=float(self.terminal_cost), out={"xN": sl["xN"], "cost": sl["cost"], "terminated": sl["term"]},
                  actions=(seed, um))
        done = torch.cuda.Event()
        done.record(comp)
        with torch.cuda.stream(st["d2h"]):
            st["d2h"].wait_event(done)
            if cost_out is not None:
                cost_out.copy_(sl["cost"], non_blocking=True)
            if xN_out is not None:
                xN_out.copy_(sl["xN"], non_blocking=True)
            # the queue kernels' sticky status words travel with the results (check_queues() would synchronise with the
            # stream, i.e. with the NEXT step's kernel, and the host could not submit the step after it in time)
            words = L.queue_status_words()
            if "status" not in sl or sl["status"].numel() != len(words):
                sl["status"] = torch.zeros(max(len(words), 1), dtype=torch.int32).pin_memory()
            for i, w in enumerate(words):
                sl["status"][i:i + 1].copy_(w.reshape(1), non_blocking=True)
            back = torch.cuda.Event()
            back.record(st["d2h"])
        sl["free"] = back
        handle = _HostRollout(back, L, {"cost": cost_out, "xN": xN_out}, sl["status"])
        return handle.wait() if wait else handle

    def terminate(self, x):
        """True where `x` (`[n]` or `[B, n]`) is a terminal state (reference: environments.py:278-288)."""
        x = np.asarray(x, dtype=float)
        out = np.zeros(x.shape[:-1], dtype=bool)
        for i, lim in self._term_limits():
            out = out | (np.abs(x[..., i]) > lim)
        return out if out.ndim else bool(out)

    def observe(self, x):
        return observation(x, self.xIsAngle)

    @property
    def o(self):
        return self.observe(self.x)

    @property
    def o_(self):
        """Observation before the last step.  The reference keeps `o_` / `o` as attributes that `step` shifts and `reset`
        does not touch (environments.py:61-75): the first transition after a reset therefore carries the LAST observation
        of the previous episode as `o_` (MBRL stores it, mbrl.py:228-234).  Kept: the data sets must equal the reference's."""
        prev = self.__dict__.get("_obs_prev_state")
        if prev is None:
            return self.observe(self.x_)
        prev = np.asarray(prev)
        return self.observe(prev if self.batched else prev.reshape(-1))


class _HostRollout(object):
    """Handle of a non-blocking `StateSpaceModel.rollout_host_random(..., wait=False)` call."""

    def __init__(self, event, library, out, status):
        self._event, self._lib, self._out, self._status = event, library, out, status

    def wait(self):
        """Blocks until the call's outputs are in host memory; returns {"cost", "xN"} (the caller's pinned tensors)."""
        self._event.synchronize()
        if int(self._status.sum()) != 0:
            self._lib.check_queues()  # raises, and resets the words
        return self._out


class LearnedStateSpaceModel(StateSpaceModel):
    """Environment whose dynamics are a learned MLP: the reference's `MBRL.nn_environment =
    StateSpaceModel(self.ode, environment.cost, x0, uDim, dt)` with `MBRL.ode = [x[n/2:], 1/dt * nn_dynamics.ode(x, u)]`
    (pygent/algorithms/mbrl.py:76, :125-130), integrated with forward Euler (`fast_step`; the planner is built with
    fastForward=True, mbrl.py:81).

    `net` is a `pygent_b200.nn_models.NNDynamics` (sparse_dyn: dxDim = xDim / 2).  Rollouts and linearisation run in
    the learned-dynamics library (include/pygent_b200_mlp.h); the problem library compiled from `cost` supplies the
    cost, its expansion and the iLQR kernels.  precision "f16x3" (default) puts the 500 x 500 layer on the tensor cores at
    the reference's fp32 precision (fp16 hi/lo split operands, ~1e-6 per evaluation), "fp32" is the CUDA-core FFMA path,
    "bf16" the lower-precision tensor-core path (2e-2 per evaluation); linearization "fd" is the reference's central
    differences (helpers.py:130-148), "analytic" the exact chain rule on the tensor cores (f16x3 / bf16)."""

    def __init__(self, net, cost, x0, uDim, dt, terminal_cost=0., precision="f16x3", linearization=None, device=None):
        n = int(np.array(x0() if callable(x0) else x0).shape[-1])
        if net.xDim != n or net.uDim != uDim or net.dxDim * 2 != n:
            raise ValueError("network dimensions (x %d, u %d, dx %d) do not match the environment (n %d, m %d, sparse_dyn)"
                             % (net.xDim, net.uDim, net.dxDim, n, uDim))
        xs = sp.symbols("x1:%d" % (n + 1))
        us = sp.symbols("u1:%d" % (uDim + 1))
        # structural stand-in for the compiled problem: position rates are the velocities, the learned part is NOT
        # in this library (its rollout kernels are never used for this environment)
        f = list(xs[n // 2:]) + [sp.Integer(0)] * (n // 2)
        system = SystemModel("learned_n%d_m%d" % (n, uDim), f, xs, us, False)
        super(LearnedStateSpaceModel, self).__init__(system, cost, x0, uDim, dt, terminal_cost=terminal_cost, substeps=1,
                                                     device=device, dtype=torch.float64)
        self.net = net
        self.precision = precision
        self.linearization = linearization or ("analytic" if precision in ("bf16", "f16x3") else "fd")
        if self.linearization == "analytic" and precision not in ("bf16", "f16x3"):
            raise ValueError("the analytic Jacobian chain runs on the tensor-core paths (precision='f16x3' or 'bf16'); use linearization='fd'")
        self.learned = True

    @property
    def dynamics(self):
        """Packed device copy of the network (rebuilt after `net.invalidate()`)."""
        return self.net.device_model(self.device)

    def ode(self, t, x, u):
        """MBRL.ode (mbrl.py:125-130) for one state or a batch of states, evaluated on the GPU."""
        x, u = np.asarray(x, dtype=float), np.asarray(u, dtype=float)
        rhs = self.dynamics.ode(np.atleast_2d(x), np.atleast_2d(u), self.dt, precision=self.precision)
        return rhs[0] if x.ndim == 1 else rhs

    def _advance(self, args, integrator):
        if len(args) == 2:
            u, dt = args
        elif len(args) == 1:
            u, dt = args[0], self.dt
        else:
            raise TypeError("step(u[, dt])")
        if abs(dt - self.dt) > 1e-15:
            raise ValueError("the learned model predicts one control interval dt = %g (mbrl.py:126)" % self.dt)
        x = self._state()
        ud = self._to_dev(u, self.uDim).reshape(1, self.uDim, self.batch)
        X, Uo = self._open_loop(x, ud)
        xn = X[0, 1].contiguous()
        c = self.library().traj_cost(X, Uo, dt=dt, t0=float(self.tt[-1]), terminal_cost=float(self.terminal_cost))
        self._xd_prev, self._xd = x, xn
        self._hist.append(xn.detach().to("cpu", torch.float64).numpy().T.copy())
        self.tt.extend([self.tt[-1] + dt])
        term = np.asarray(self.terminate(self._hist[-1]))
        cost = c["cost"][0].to("cpu", torch.float64).numpy()
        # traj_cost adds fcost(x_T)*dt, which is zero for the plain environment library
        if self.batched:
            self.terminated = term
            return cost
        self.terminated = bool(term[0]) if term.ndim else bool(term)
        return float(cost[0])

    def _open_loop(self, x, Ud):
        """Euler rollout along given controls Ud [T, m, B]: (X [1,T+1,n,B], U [1,T,m,B])."""
        X = self.dynamics.rollout_device(x, self.dt, U=Ud, precision=self.precision)
        return X[None], Ud[None]

    def step(self, *args):
        """The learned model has one integrator: forward Euler over one control interval."""
        return self._advance(args, lib.INTEG_EULER)

    fast_step = step

    def rollout(self, U, history=True, fast=True, add_final_cost=False, fcost_expr=None):
        x = self._state()
        if torch.is_tensor(U) and U.is_cuda:
            Ud = U
        else:
            U = np.asarray(U, dtype=float)
            if U.ndim == 2:
                U = np.broadcast_to(U[None], (self.batch,) + U.shape)
            Ud = torch.as_tensor(np.ascontiguousarray(U.transpose(1, 2, 0)), dtype=self.dtype, device=self.device)
        X, Uo = self._open_loop(x, Ud)
        c = self.library(fcost_expr).traj_cost(X, Uo, dt=self.dt, t0=float(self.tt[-1]), terminal_cost=float(self.terminal_cost))
        T = Ud.shape[0]
        self._xd_prev, self._xd = x, X[0, T].contiguous()
        Xh = X[0].detach().to("cpu", torch.float64).numpy()
        self._hist.extend([Xh[k].T.copy() for k in range(1, T + 1)] if history else [Xh[T].T.copy()])
        for _ in range(T):
            self.tt.append(self.tt[-1] + self.dt)
        return {"X": X[0], "xN": self._xd, "cost": c["cost"][0], "terminated": c["terminated"][0]}


def _finish(env):
    env.oDim = int(env.xDim + np.sum(env.xIsAngle))


class Pendulum(StateSpaceModel):
    """reference: environments.py:328-357"""

    def __init__(self, cost, x0, dt, **kw):
        super(Pendulum, self).__init__(get_system("pendulum"), cost, x0, 1, dt, **kw)
        self.xIsAngle = [True, False]
        self.uMax = 3.5 * np.ones(1)
        _finish(self)

    def _term_limits(self):
        return [(1, 10.0), (0, 8 * np.pi)]


class AngularPendulum(StateSpaceModel):
    """reference: environments.py:394-423"""

    def __init__(self, cost, x0, dt, **kw):
        super(AngularPendulum, self).__init__(get_system("angular_pendulum"), cost, x0, 1, dt, **kw)
        self.xIsAngle = [False, False, False]
        self.uMax = 3.5 * np.ones(1)
        _finish(self)

    def _term_limits(self):
        return [(2, 10.0)]


class CartPole(StateSpaceModel):
    """reference: environments.py:461-482 (`linearized=True`: input = cart acceleration)"""

    def __init__(self, cost, x0, dt, linearized=True, **kw):
        super(CartPole, self).__init__(get_system("cart_pole_lin" if linearized else "cart_pole"), cost, x0, 1, dt, **kw)
        self.xIsAngle = [False, True, False, False]
        self.uMax = (10 if linearized else 100) * np.ones(1)
        self.x1Max = 1.2
        self.x3Max = 4.0
        _finish(self)

    def _term_limits(self):
        return [(0, self.x1Max), (2, self.x3Max)]


class CartPoleDoubleSerial(StateSpaceModel):
    """reference: environments.py:523-551"""

    def __init__(self, cost, x0, dt, linearized=True, task='swing_up', **kw):
        name = "cart_pole_double_serial_lin" if linearized else "cart_pole_double_serial"
        super(CartPoleDoubleSerial, self).__init__(get_system(name), cost, x0, 1, dt, **kw)
        self.xIsAngle = [False, True, True, False, False, False]
        self.uMax = (15 if linearized else 100) * np.ones(1)
        self.task = task
        self.x1Max = 1.5
        _finish(self)

    def _term_limits(self):
        if self.task == 'swing_up':
            return [(0, self.x1Max), (4, 25.0), (5, 25.0)]
        if self.task == 'balance':
            return [(0, self.x1Max), (1, 1.0), (2, 1.0), (4, 25.0), (5, 25.0)]
        return []


class CartPoleDoubleParallel(StateSpaceModel):
    """reference: environments.py:596-615"""

    def __init__(self, cost, x0, dt, linearized=True, **kw):
        name = "cart_pole_double_parallel_lin" if linearized else "cart_pole_double_parallel"
        super(CartPoleDoubleParallel, self).__init__(get_system(name), cost, x0, 1, dt, **kw)
        self.xIsAngle = [False, True, True, False, False, False]
        self.uMax = (25 if linearized else 100) * np.ones(1)
        self.x1Max = 1.5
        _finish(self)

    def _term_limits(self):
        return [(0, self.x1Max), (4, 25.0), (5, 25.0)]


class CartPoleTriple(StateSpaceModel):
    """reference: environments.py:660-679 (only the linearised model is shipped)"""

    def __init__(self, cost, x0, dt, linearized=True, **kw):
        if not linearized:
            raise NotImplementedError("the reference ships only cart_pole_triple_lin (c_files listing)")
        super(CartPoleTriple, self).__init__(get_system("cart_pole_triple_lin"), cost, x0, 1, dt, **kw)
        self.xIsAngle = [False, True, True, True, False, False, False, False]
        self.uMax = 20 * np.ones(1)
        self.x1Max = 1.5
        _finish(self)

    def _term_limits(self):
        return [(0, self.x1Max)]


class Car(StateSpaceModel):
    """reference: environments.py:730-761 (m = 2)"""

    def __init__(self, cost, x0, dt, **kw):
        super(Car, self).__init__(get_system("car"), cost, x0, 2, dt, **kw)
        self.xIsAngle = [False, False, True, False]
        self.uMax = np.array([.5, 2.])
        _finish(self)

    def _term_limits(self):
        return [(0, 5.0), (1, 5.0)]


class Acrobot(StateSpaceModel):
    """reference: environments.py:763-782"""

    def __init__(self, cost, x0, dt, linearized=True, **kw):
        super(Acrobot, self).__init__(get_system("acrobot_lin" if linearized else "acrobot"), cost, x0, 1, dt, **kw)
        self.xIsAngle = [True, True, False, False]
        self.uMax = (50 if linearized else 5) * np.ones(1)
        _finish(self)

    def _term_limits(self):
        return [(2, 50.0), (3, 50.0)]


class MarBot(StateSpaceModel):
    """reference: environments.py:826-868"""

    def __init__(self, cost, x0, dt, **kw):
        super(MarBot, self).__init__(get_system("marbot"), cost, x0, 1, dt, **kw)
        self.xIsAngle = [False, True, False, False]
        self.uMax = 2 * np.ones(1)
        _finish(self)

    def _term_limits(self):
        return [(0, 1.0)]
