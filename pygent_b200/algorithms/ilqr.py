"""Batched iLQR on the GPU behind pygent's `iLQR` API.

Mirrors `pygent/algorithms/ilqr.py` (reference): constructor kwargs (:56-75), `init_trajectory`
(:501-516), `forward_pass` (:187-231), `backward_pass` (:233-346), `run_optim` (:399-499),
`linearization` (:518-533), `cost_lin` (:597-610), `run` (:348-351), `reset` (:164-171), `save`
(:646-656).  Every trajectory of the environment's batch is optimised independently, all at once:

    per iteration   K3 pg_ilqr_backward   one thread per trajectory
                    K2 pg_ilqr_forward    one thread per (trajectory, alpha): all 11 alphas at once
                    K4 pg_ilqr_select     the reference's serial (or `parallel=True`) line-search
                                          bookkeeping, accept/reject, mu schedule, stopping tests
                    K2' pg_ilqr_commit    re-run of the winning alpha -> new nominal trajectory

Evaluating all alphas speculatively changes nothing: K4 replays the reference's control flow on the
11 costs (first alpha with z > 0 stops a serial search; argmin over the alphas tried so far; dcost
of the last alpha tried feeds the tolFun test -- SURVEY F7).  Trajectories stop individually
(`status`), the loop ends when none is running.

Learned dynamics (`environments.LearnedStateSpaceModel`, the reference's MBRL planner, mbrl.py:76-93): the same loop
with the dynamics' own library in place of the generated ODE --

    per iteration   pg_mlp_linearize      Jacobians along the nominal trajectory (analytic chain on the tensor
                                          cores, or the reference's central differences)
                    pg_ilqr_backward_ext  Riccati sweep on those Jacobians
                    pg_mlp_forward        all 11 closed-loop Euler rollouts through the network
                    pg_traj_cost          their costs;  K4 as above;  pg_ilqr_commit_copy instead of a re-run

Single-plant environments return arrays shaped like the reference's (`xx [T+1,n]`, `uu [T,m]`,
`KK [T,m,n]`, `kk [T,m,1]`); batched ones gain a leading B axis.
"""
import os
import pickle
import time

import numpy as np
import sympy as sp
import torch

from .. import lib, tracing
from ..agents import FeedBack
from .core import Algorithm


class iLQR(Algorithm):
    def __init__(self, environment, t, dt,
                 maxIters=500,
                 tolGrad=1e-4,
                 tolFun=1e-7,
                 fastForward=False,
                 path=None,
                 fcost=None,
                 constrained=False,
                 save_interval=10,
                 printing=True,
                 log_data=False,
                 dataset_size=1e6,
                 regType=2,
                 finite_diff=False,
                 file_prefix='',
                 init=True,
                 reset_mu=True,
                 saving=True,
                 parallel=False,
                 final_backpass=True,
                 check_every=4, use_graph=None,
                 cache_terminal_value=False):
        self.uDim = environment.uDim
        self.xDim = environment.xDim
        self.saving = path is not None
        self.path = path
        if self.saving:
            for sub in ("", "plots/", "animations/", "data/", "c_files/"):
                os.makedirs(path + sub, exist_ok=True)
        self.fastForward = fastForward
        batch = environment.batch if environment.batched else None
        super(iLQR, self).__init__(environment, FeedBack(None, self.uDim, batch), t, dt)
        self.xIsAngle = self.environment.xIsAngle
        self.finite_diff = finite_diff
        self.final_backpass = final_backpass
        # final backward pass: the terminal value P depends on the final-cost equilibrium only; with this flag it is
        # solved once (host: scipy minimise + discrete Riccati equation) and reused, which removes the host round trip
        # from every later run_optim / MPC control step.  Off by default: the reference re-solves it every time.
        self.cache_terminal_value = cache_terminal_value
        self._P_cache = None
        # final cost expression: fcost(x) or fcost(x, mod) called on symbols (ilqr.py:119-125, :552)
        xs = list(environment.system.xs)
        if fcost is None:
            # the reference's default, cost_fnc(x, 0, t, mod) = cost*dt, then scaled by dt again
            c = environment.cost_expression()
            zero_u = {u: 0 for u in environment.system.us}
            self._fcost_expr = sp.sympify(c).subs(zero_u).subs(sp.Symbol("t"), t) * dt
        else:
            self._fcost_expr = tracing.trace(tracing.bind_fcost(fcost), xs, sp)
        self.learned = bool(getattr(environment, "learned", False))
        if self.learned and not fastForward:
            raise NotImplementedError("a learned model predicts one Euler step per control interval: construct the "
                                      "planner with fastForward=True, as MBRL does (mbrl.py:81)")
        self.lib = environment.library(self._fcost_expr)
        if not self.lib.info.has_cost_derivs:
            raise ValueError("iLQR needs a running cost c(x, u[, t]) that does not depend on the next state "
                             "(the reference evaluates it with x' = None, ilqr.py:183)")
        self.integrator = lib.INTEG_EULER if fastForward else lib.INTEG_RK4
        self.substeps = environment.substeps
        self.lin_mode = 2 if self.learned else (0 if self.lib.has_ab else 1)
        if finite_diff:
            # the cost expansion by central differences (helpers.py:41-128 via ilqr.py:240-242, :598-603): same kernels,
            # PG_LIN_COST_FD or-ed into the linearisation mode
            if environment.dtype != torch.float64:
                raise ValueError("finite_diff=True needs fp64: second differences with eps = 1e-3 are noise in fp32")
            self.lin_mode |= lib.LIN_COST_FD
        self.dtype = environment.dtype
        self.B = environment.batch
        self.batched = environment.batched
        self.reset_mu = reset_mu
        self.init = init
        self.printing = printing
        self.file_prefix = file_prefix
        self.check_every = max(1, int(check_every))
        self.use_graph = use_graph  # None: replay iterations from a CUDA graph when possible (see _iterate_until_done)

        # algorithm parameters (ilqr.py:141-162)
        self.success_fw = False
        self.max_iters = maxIters
        self.mu_min = 1e-6
        self.mu_max = 1e10
        self.mu_d0 = 1.6
        self.mu0 = 1.e-6
        self.alphas = 10 ** np.linspace(0, -3, 11)
        self.zmin = 0.
        self.tolGrad = tolGrad
        self.tolFun = tolFun
        self.constrained = constrained
        self.regType = regType
        self.save_interval = save_interval
        self.parallel = parallel
        self.tt = np.arange(0, self.t, self.dt)
        self.iterations = 0
        self._alloc()
        if init:
            self.init_trajectory()

    # -- device state -------------------------------------------------------------------------------
    @property
    def lims(self):
        return np.asarray(self.environment.uMax, dtype=float)  # read at use: examples rescale env.uMax

    def _alloc(self):
        T, n, m, B = self.steps, self.xDim, self.uDim, self.B
        dev, dt_ = self.environment.device, self.dtype
        z = lambda *s: torch.zeros(s, dtype=dt_, device=dev)
        self._xbar, self._ubar = z(T + 1, n, B), z(T, m, B)
        self._KK, self._kk = z(T, m, n, B), z(T, m, B)            # accepted gains (self.KK / self.kk)
        self._KKc, self._kkc = z(T, m, n, B), z(T, m, B)          # candidate gains of the current backward pass
        self._x0 = self.environment._to_dev(self.environment._x0_host, n)
        u8 = lambda *s: torch.zeros(s, dtype=torch.uint8, device=dev)
        i32 = lambda *s: torch.zeros(s, dtype=torch.int32, device=dev)
        # Everything a new solve re-arms lives in ONE block with a template beside it: status, iters, the first work list and
        # its count, active, success_gradient (and, in a second block, mu / mu_d, which reset_mu=False keeps) -- a solve
        # starts with one or two device copies instead of eight fill kernels (a receding-horizon control step is ~20 launches)
        Bp = (B + 15) // 16 * 16
        self._arm = torch.zeros(4 * (3 * Bp + 16) + 2 * Bp, dtype=torch.uint8, device=dev)
        i32v = lambda off, cnt: self._arm[4 * off:4 * (off + cnt)].view(torch.int32)
        status_, iters_, idx0_, cnt0_ = i32v(0, B), i32v(Bp, B), i32v(2 * Bp, B), i32v(3 * Bp, 1)
        active_ = self._arm[4 * (3 * Bp + 16):4 * (3 * Bp + 16) + B]
        sgrad_ = self._arm[4 * (3 * Bp + 16) + Bp:4 * (3 * Bp + 16) + Bp + B]
        idx0_.copy_(torch.arange(B, dtype=torch.int32, device=dev))
        cnt0_.fill_(B)
        active_.fill_(1)
        self._arm_template = self._arm.clone()
        self._mu_block = z(2, B)
        self._mu_block[0] = self.mu0
        self._mu_block[1] = 1.0
        self._mu_template = self._mu_block.clone()
        self._mu_template_mu0 = float(self.mu0)
        self._st = {"cost": z(B), "mu": self._mu_block[0], "mu_d": self._mu_block[1], "dcost": z(B),
                    "success_gradient": sgrad_, "status": status_, "iters": iters_, "alpha_best": z(B) + 1.0,
                    "accept": u8(B), "active": active_}
        A = len(self.alphas)
        self._bw = {"KK": self._KKc, "kk": self._kkc, "dV": z(2, B), "gnorm": z(B), "success": u8(B)}
        self._fw = {"cost": z(A, B), "diverged": u8(A, B), "terminated": u8(A, B), "xN": z(A, n, B)}
        # chunk-parallel commit: the candidates' states at the 20-step chunk boundaries
        self._n_snap = 0 if self.learned else self.lib.num_snapshots(T)
        if self._n_snap > 0:
            self._fw["xS"] = z(A, self._n_snap, n, B)
        self._last_tried = i32(B) - 1   # last alpha evaluated per trajectory (MPC plan shift)
        # compacted work lists (ping-pong): trajectories still running, written by the select kernel
        self._idx = [idx0_, i32(B)]
        self._cnt = [cnt0_, i32(1)]
        self._cur = 0
        self._n_active = self._cnt[1]
        self._have_gains = False
        # serial line search in two launches: work lists "backward pass succeeded" and "no success among the first alphas"
        self.first_alphas = 4
        # live trajectories below which the solve switches to its tail form: the fused per-trajectory solver once they all fit
        # on the GPU at one warp each (lib.solve_capacity(): 1,184 on a B200), else iterate()'s one-launch line search
        cap = self.lib.solve_capacity() if (self.environment.device.type == "cuda" and not self.learned) else 0
        self.tail_threshold = int(os.environ.get("PG_ILQR_TAIL", str(cap if cap > 0 else 2048)))
        self.final_on_device = None  # final backward pass: None = per-trajectory on the device when B > 32, host (scipy) otherwise
        self.final_status = None
        self._trace = None  # list: (iteration, live trajectories, host time) per readback, for tools/time_fused_tail.py
        self._round_trace = []
        self.tail_round = int(os.environ.get("PG_ILQR_TAIL_ROUND", "32"))  # iterations per launch of the fused tail solver
        self.fused_tail = os.environ.get("PG_ILQR_FUSED_TAIL", "1") != "0"  # ... or, better, the fused per-trajectory solver finishes the solve
        self._tail = False
        self._ls_small = lib.split_search_threshold(self._x0.device)
        self._ls_idx = [i32(B), i32(B)]
        self._ls_cnt = i32(2)
        if self.learned:  # Jacobians along the nominal trajectory and the candidate trajectories of the line search
            self._lin = {"A": z(T, n, n, B), "B": z(T, n, m, B)}
            self._cand = {"X": z(A, T + 1, n, B), "U": z(A, T, m, B)}

    def _host(self, t, batch_axis_last=True):
        t = t.detach()
        if t.is_cuda and t.dtype == torch.float64 and t.numel() >= (1 << 17):
            # large results (the nominal trajectories of a 16,384-trajectory batch are 66 MB) go through page-locked memory:
            # ~50 GB/s instead of the ~3 GB/s of a copy into pageable memory.  The array owns its buffer (torch's caching
            # host allocator hands the block out again once the array is gone).
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t, non_blocking=True)
            torch.cuda.current_stream(t.device).synchronize()
            a = h.numpy()
        else:
            a = t.to("cpu", torch.float64).numpy()
        a = np.moveaxis(a, -1, 0)  # [B, ...]
        return a if self.batched else a[0]

    # reference-shaped views ------------------------------------------------------------------------
    @property
    def xx(self):
        return self._host(self._xbar)

    @property
    def uu(self):
        return self._host(self._ubar)

    @property
    def KK(self):
        return self._host(self._KK) if self._have_gains else []

    @property
    def kk(self):
        return self._host(self._kk)[..., None] if self._have_gains else []

    @property
    def cost(self):
        c = self._st["cost"].to("cpu", torch.float64).numpy()
        return c if self.batched else float(c[0])

    @property
    def mu(self):
        v = self._st["mu"].to("cpu", torch.float64).numpy()
        return v if self.batched else float(v[0])

    @property
    def current_alpha(self):
        v = self._st["alpha_best"].to("cpu", torch.float64).numpy()
        return v if self.batched else float(v[0])

    @property
    def status(self):
        v = self._st["status"].cpu().numpy()
        return v if self.batched else int(v[0])

    @property
    def iters(self):
        v = self._st["iters"].cpu().numpy()
        return v if self.batched else int(v[0])

    def _ulim(self):
        return list(self.lims) if self.constrained else None

    # -- the reference's methods ---------------------------------------------------------------------
    def init_trajectory(self):
        """Initial trajectory with u = 0 (ilqr.py:501-516): cost = sum c*dt + fcost(x_N)*dt."""
        self._x0 = self.environment._to_dev(self.environment._x0_host, self.xDim)
        self._ubar.zero_()
        self._have_nominal = True
        if self.learned:
            env = self.environment
            X = env.dynamics.rollout_device(self._x0, self.dt, T=self.steps, precision=env.precision)
            self._xbar.copy_(X)
            c = self.lib.traj_cost(X[None], self._ubar[None], dt=self.dt, terminal_cost=float(env.terminal_cost))
            self._st["cost"].copy_(c["cost"][0])
            self.agent.reset()
            return
        self.lib.rollout(self._x0, None, T=self.steps, S=self.substeps, dt=self.dt, integrator=self.integrator,
                         terminal_cost=float(self.environment.terminal_cost), add_final_cost=True,
                         out={"X": self._xbar, "cost": self._st["cost"]})
        self.agent.reset()

    def reset(self):
        self._have_gains = False
        self._KK.zero_()
        self._kk.zero_()
        self._arm_state(True)
        if self.init or not getattr(self, "_have_nominal", False):  # the reference also initialises when it has no plan yet
            self.init_trajectory()

    def cost_fnc(self, x, u, t, mod):
        return self.environment.cost(x, u, None, t, mod) * self.dt

    def _gains_to_dev(self, KK, kk):
        T, n, m, B = self.steps, self.xDim, self.uDim, self.B
        KK = np.asarray(KK, dtype=float).reshape((B, -1, m, n) if self.batched else (1, -1, m, n))
        kk = np.asarray(kk, dtype=float).reshape((B, -1, m) if self.batched else (1, -1, m))
        dev = self.environment.device
        return (torch.as_tensor(np.ascontiguousarray(np.moveaxis(KK, 0, -1)), dtype=self.dtype, device=dev),
                torch.as_tensor(np.ascontiguousarray(np.moveaxis(kk, 0, -1)), dtype=self.dtype, device=dev))

    def forward_pass(self, alpha, KK, kk, optim=True):
        """Closed-loop rollout for one alpha from environment.x0 (ilqr.py:187-231).
        Returns xx, uu, cost, terminated (host arrays, reference shapes)."""
        if len(KK) == 0:
            raise ValueError("forward_pass needs gains (run backward_pass or run_optim first)")
        KKd, kkd = (KK, kk) if torch.is_tensor(KK) else self._gains_to_dev(KK, kk)
        r = self._forward_candidates([float(alpha)], KKd, kkd)
        xx, uu = self._host(r["X"][0]), self._host(r["U"][0])
        cost = r["cost"][0].to("cpu", torch.float64).numpy()
        term = r["terminated"][0].cpu().numpy().astype(bool)
        self.agent.set_history(uu)
        return (xx, uu, cost, term) if self.batched else (xx, uu, float(cost[0]), bool(term[0]))

    def _forward_candidates(self, alphas, KKd, kkd, out=None, cand=None, work=None):
        """Closed-loop rollouts for `alphas` from x0 with trajectories: dict(X, U, cost, diverged, terminated)."""
        env = self.environment
        if self.learned:
            X, U = env.dynamics.forward_device(self._x0, self.dt, alphas, (KKd, kkd, self._xbar, self._ubar), ulim=self._ulim(),
                                               precision=env.precision, out=cand, work=work)
            r = self.lib.traj_cost(X, U, dt=self.dt, terminal_cost=float(env.terminal_cost), out=out, work=work)
            r.update(X=X, U=U)
            return r
        return self.lib.ilqr_forward(self._x0, self._xbar, self._ubar, KKd, kkd, alphas=alphas, S=self.substeps, dt=self.dt,
                                     integrator=self.integrator, ulim=self._ulim(),
                                     terminal_cost=float(env.terminal_cost), want_traj=True, out=out, work=work)

    def _backward(self, work=None, Vxx_T=None):
        """One Riccati sweep into the candidate gains.  Vxx_T: the final backward pass (mu = 0, unconstrained)."""
        final = Vxx_T is not None
        mu = torch.zeros_like(self._st["mu"]) if final else self._st["mu"]
        ulim = None if final else self._ulim()
        if self.learned:
            env = self.environment
            env.dynamics.linearize_device(self._xbar, self._ubar, self.dt, mode=env.linearization, precision=env.precision,
                                          out=self._lin, work=work)
            return self.lib.ilqr_backward_ext(self._xbar, self._ubar, mu, self._lin["A"], self._lin["B"], dt=self.dt,
                                              reg_type=self.regType, ulim=ulim, out=self._bw, work=work, Vxx_T=Vxx_T,
                                              cost_fd=self.finite_diff)
        return self.lib.ilqr_backward(self._xbar, self._ubar, mu, dt=self.dt, reg_type=self.regType,
                                      lin_mode=self.lin_mode, ulim=ulim, out=self._bw, work=work, Vxx_T=Vxx_T)

    # -- final backward pass (ilqr.py:246-260, :487-494) ----------------------------------------------------------------
    def _terminal_value(self):
        """P [n, n, B]: solution of the discrete Riccati equation at the equilibrium of the final cost, computed on the
        host exactly as the reference does: `scipy.optimize.minimize(fcost, x_N)`, linearisation and cost expansion at
        (x*, u = 0, t = tt[-1]), `scipy.linalg.solve_discrete_are(Ad, Bd, Cxx, Cuu, s=Cxu)`.  Per trajectory for small
        batches; for B > 32 the equilibrium search starts from trajectory 0's terminal state and the result is shared
        (the final cost has one minimiser in every shipped example)."""
        from scipy import linalg, optimize
        xs = list(self.environment.system.xs)
        f_l = sp.lambdify(xs, self._fcost_expr, modules="numpy")
        f = lambda x: float(f_l(*x))
        xN = np.atleast_2d(self._host(self._xbar[self.steps]))
        reps = range(self.B) if self.B <= 32 else [0]
        x_eq = np.array([optimize.minimize(f, xN[b]).x for b in reps])
        u0 = np.zeros((len(x_eq), self.uDim))
        Ad, Bd, _ = self.linearization(x_eq, u0)
        Cxx, Cuu, Cxu, _, _ = self.cost_lin(x_eq, u0, float(self.tt[-1]))
        Ad, Bd, Cxx, Cuu, Cxu = (np.asarray(a).reshape((len(x_eq),) + np.asarray(a).shape[-2:]) for a in (Ad, Bd, Cxx, Cuu, Cxu))
        P = np.array([linalg.solve_discrete_are(Ad[i], Bd[i], Cxx[i], Cuu[i], s=Cxu[i]) for i in range(len(x_eq))])
        if len(P) != self.B:
            P = np.broadcast_to(P[0], (self.B,) + P.shape[1:])
        return torch.as_tensor(np.ascontiguousarray(np.moveaxis(P, 0, -1)), dtype=self.dtype, device=self.environment.device)

    def _terminal_value_device(self):
        """P [n, n, B] per trajectory on the device (lib.terminal_value: Newton on the final cost + the doubling algorithm
        for the Riccati equation); (P, ok [B] bool) or None where the device route does not apply (learned dynamics, fp32,
        `finite_diff=True`)."""
        if self.learned or self.finite_diff or self.dtype != torch.float64:
            return None
        r = self.lib.terminal_value(self._xbar, self.dt, float(self.tt[-1]), lin_mode=self.lin_mode)
        if r is None:
            return None
        return r["P"], r["ok"].bool()

    def _final_backpass(self):
        """`self.mu = 0; backward_pass(final=True); if success: self.KK, self.kk = KK, kk` for every trajectory.
        Batches larger than 32 solve their equilibrium / Riccati problems on the device, one per trajectory
        (`final_on_device`; small batches keep the host route, which calls the reference's own scipy functions)."""
        converged = None
        if self._P_cache is not None:
            P = self._P_cache
        else:
            dev_route = self._terminal_value_device() if (self.final_on_device if self.final_on_device is not None else self.B > 32) else None
            if dev_route is not None:
                P, converged = dev_route
            else:
                try:
                    P = self._terminal_value()
                except (np.linalg.LinAlgError, ValueError) as e:
                    # no stabilising solution: the reference raises here (scipy.linalg.solve_discrete_are); so do we
                    self.final_status = "failed: %s" % e
                    raise
        if self.cache_terminal_value:
            self._P_cache = P
        r = self._backward(Vxx_T=P)
        ok = r["success"].bool()
        if converged is not None:
            ok = ok & converged  # no terminal value, no final gains for that trajectory (its iLQR gains stay)
        self.final_status = "ok"
        self._KK.copy_(torch.where(ok, self._KKc, self._KK))
        self._kk.copy_(torch.where(ok, self._kkc, self._kk))
        self._st["mu"].zero_()
        self._have_gains = True

    def backward_pass(self, final=False):
        """Riccati sweep along (xx, uu) (ilqr.py:233-346).  Returns V1, V2, success, KK, kk where V1, V2
        are the SUMS of the reference's per-step lists (only their sums are ever used, ilqr.py:442)."""
        r = self._backward(Vxx_T=self._terminal_value() if final else None)
        dV = r["dV"].to("cpu", torch.float64).numpy()
        ok = r["success"].cpu().numpy().astype(bool)
        KK, kk = self._host(r["KK"]), self._host(r["kk"])[..., None]
        if self.batched:
            return dV[0], dV[1], ok, KK, kk
        return float(dV[0, 0]), float(dV[1, 0]), bool(ok[0]), KK, kk

    def linearization(self, x, u):
        """Ad = I + A dt, Bd = B dt, fd = 0 (ilqr.py:518-533) at one point or a batch of points."""
        x, u = np.atleast_2d(np.asarray(x, dtype=float)), np.atleast_2d(np.asarray(u, dtype=float))
        dev = self.environment.device
        xd = torch.as_tensor(x.T.copy(), dtype=self.dtype, device=dev)
        ud = torch.as_tensor(u.T.copy(), dtype=self.dtype, device=dev)
        if self.learned:
            env = self.environment
            Ad_, Bd_ = env.dynamics.linearize_device(torch.stack([xd, xd]), ud[None], self.dt, mode=env.linearization,
                                                     precision=env.precision)
            r = {"A": Ad_[0], "B": Bd_[0]}
        else:
            r = self.lib.eval(xd, ud, lin_mode=self.lin_mode & 3, what=("A", "B"))
        A = np.moveaxis(r["A"].to("cpu", torch.float64).numpy(), -1, 0)
        Bm = np.moveaxis(r["B"].to("cpu", torch.float64).numpy(), -1, 0)
        Ad = A * self.dt + np.eye(self.xDim)
        Bd = Bm * self.dt
        fd = np.zeros((len(x), self.xDim, 1))
        return (Ad[0], Bd[0], fd[0]) if len(x) == 1 else (Ad, Bd, fd)

    def cost_lin(self, x, u, t):
        """Second-order expansion of c*dt: Cxx, Cuu, Cxu, cx, cu (ilqr.py:597-610)."""
        x, u = np.atleast_2d(np.asarray(x, dtype=float)), np.atleast_2d(np.asarray(u, dtype=float))
        dev = self.environment.device
        xd = torch.as_tensor(x.T.copy(), dtype=self.dtype, device=dev)
        ud = torch.as_tensor(u.T.copy(), dtype=self.dtype, device=dev)
        td = torch.as_tensor(np.broadcast_to(np.asarray(t, dtype=float), (len(x),)).copy(), dtype=self.dtype, device=dev)
        r = self.lib.eval(xd, ud, td, lin_mode=self.lin_mode & ~3 if self.finite_diff else None,
                          what=("cx", "cu", "Cxx", "Cuu", "Cxu"))
        g = lambda k: np.moveaxis(r[k].to("cpu", torch.float64).numpy(), -1, 0) * self.dt
        out = (g("Cxx"), g("Cuu"), g("Cxu"), g("cx")[..., None], g("cu")[..., None])
        return tuple(o[0] for o in out) if len(x) == 1 else out

    def _select_params(self):
        p = lib.SelectParams()
        p.n_alpha, p.parallel, p.max_iters = len(self.alphas), int(self.parallel), int(self.max_iters)
        for i, a in enumerate(self.alphas):
            p.alphas[i] = float(a)
        p.zmin, p.tol_fun, p.tol_grad = self.zmin, self.tolFun, self.tolGrad
        p.mu_min, p.mu_max, p.mu_d0 = self.mu_min, self.mu_max, self.mu_d0
        return p

    def _rearm(self):
        """Every trajectory back on the work list, nothing else touched (tests and tools that drive `iterate()` themselves)."""
        Bp = (self.B + 15) // 16 * 16
        tmpl = self._arm_template[4 * 2 * Bp:4 * (3 * Bp + 1)].view(torch.int32)
        self._idx[0].copy_(tmpl[:self.B])
        self._cnt[0].copy_(tmpl[Bp:Bp + 1])
        self._cur = 0
        self._tail = False
        self._n_active = self._cnt[0]

    def _arm_state(self, reset_mu):
        """Start of a solve (`for _ in range(1, max_iters + 1)` starts over): status, iteration counts, stopping flags, every
        trajectory back on the work list -- one device copy of the template block; mu, mu_d with `reset_mu` -- a second."""
        self._arm.copy_(self._arm_template)
        if reset_mu:
            if float(self._mu_template_mu0) != float(self.mu0):  # mu0 is a public attribute: follow it
                self._mu_template[0] = self.mu0
                self._mu_template_mu0 = float(self.mu0)
            self._mu_block.copy_(self._mu_template)
        self._cur = 0
        self._tail = False
        self._n_active = self._cnt[0]

    def iterate(self):
        """One loop body of run_optim (ilqr.py:407-485) for every running trajectory: 4 launches, no sync.
        The kernels walk the compacted list of running trajectories; select writes the next list."""
        st = self._st
        L = self.lib
        cur, nxt = self._cur, 1 - self._cur
        work = (self._idx[cur], self._cnt[cur])
        self._backward(work=work)
        if self.learned:
            self._forward_candidates(list(self.alphas), self._KKc, self._kkc, out=self._fw, cand=self._cand, work=work)
        elif self._tail:
            # few trajectories left (the host has seen the count, one block late): every rollout is a latency-bound chain,
            # so ONE launch of the plain kernel for all alphas -- no queue hand-offs between time chunks, no second launch
            L.ilqr_forward(self._x0, self._xbar, self._ubar, self._KKc, self._kkc, alphas=self.alphas, S=self.substeps,
                           dt=self.dt, integrator=self.integrator, ulim=self._ulim(),
                           terminal_cost=float(self.environment.terminal_cost), out=self._fw, work=work, balanced=False)
        elif (not self.parallel) and 0 < self.first_alphas < len(self.alphas) and self.B > self._ls_small:
            # serial search in two launches: the first alphas for everybody whose backward pass succeeded, the others only
            # where nothing has succeeded yet -- as long as more than _ls_small trajectories are live (decided on the device)
            nf, al = self.first_alphas, [float(v) for v in self.alphas]
            self._ls_cnt.zero_()
            w0 = (self._ls_idx[0], self._ls_cnt[0:1])
            w1 = (self._ls_idx[1], self._ls_cnt[1:2])
            L.linesearch_filter(0, self._bw["success"], w0[0], w0[1], work=work)
            fw = dict(alphas=al, S=self.substeps, dt=self.dt, integrator=self.integrator, ulim=self._ulim(),
                      terminal_cost=float(self.environment.terminal_cost), out=self._fw,
                      balanced=lib.use_balanced_forward(self.B, len(al), self._ubar.shape[0], self._x0.device))
            L.ilqr_forward(self._x0, self._xbar, self._ubar, self._KKc, self._kkc, work=w0,
                           window=(1, nf, self._ls_small, w0[1]), **fw)
            L.linesearch_filter(1, self._bw["success"], w1[0], w1[1], work=w0, n_first=nf, alphas=al[:nf], zmin=self.zmin,
                                cost=st["cost"], dV=self._bw["dV"], cost_cand=self._fw["cost"], diverged=self._fw["diverged"],
                                small_count=self._ls_small)
            L.ilqr_forward(self._x0, self._xbar, self._ubar, self._KKc, self._kkc, work=w1,
                           window=(2, nf, self._ls_small, w0[1]), **fw)
        else:
            L.ilqr_forward(self._x0, self._xbar, self._ubar, self._KKc, self._kkc, alphas=self.alphas, S=self.substeps,
                           dt=self.dt, integrator=self.integrator, ulim=self._ulim(),
                           terminal_cost=float(self.environment.terminal_cost), out=self._fw, work=work)
        self._cnt[nxt].zero_()
        L.ilqr_select(self._sel, self._fw["cost"], self._fw["diverged"], self._bw["dV"], self._bw["gnorm"],
                      self._bw["success"], st, self._cnt[nxt], work=work, index_out=self._idx[nxt],
                      last_tried=self._last_tried)
        if self.learned:
            L.ilqr_commit_copy(list(self.alphas), st["alpha_best"], st["accept"], self._cand["X"], self._cand["U"], self._xbar,
                               self._ubar, self._KKc, self._kkc, self._KK, self._kk, work=work)
        elif self._n_snap > 0:
            L.ilqr_commit_chunked(list(self.alphas), self._x0, self._xbar, self._ubar, self._KKc, self._kkc, self._KK, self._kk,
                                  st["alpha_best"], st["accept"], self._fw["xS"], S=self.substeps, dt=self.dt,
                                  integrator=self.integrator, ulim=self._ulim(), work=work)
        else:
            L.ilqr_commit(self._x0, self._xbar, self._ubar, self._KKc, self._kkc, self._KK, self._kk, st["alpha_best"],
                          st["accept"], S=self.substeps, dt=self.dt, integrator=self.integrator, ulim=self._ulim(), work=work)
        self._cur = nxt
        self._n_active = self._cnt[nxt]
        self._have_gains = True

    def run_iterations(self, n):
        """`n` loop bodies of run_optim for every trajectory, device only: no host synchronisation, nothing copied
        back (the receding-horizon agent calls this once per control step)."""
        self._arm_state(self.reset_mu)
        self._sel = self._select_params()
        self._sel.max_iters = int(n)
        for _ in range(int(n)):
            self.iterate()
        if self.final_backpass:
            self._final_backpass()

    def _iterate_until_done(self):
        """Loop bodies of run_optim in blocks of `check_every` until no trajectory is running or max_iters is reached.
        The number of running trajectories comes back through pinned memory one block LATE (the next block is already
        enqueued when the host looks at it), so the device never waits for the host; the price is at most one block of
        iterations over an empty work list.  After the first block the iterations are replayed from a CUDA graph of two
        consecutive ones (the work lists ping-pong between two buffers), which takes the ~9 ctypes / torch calls per
        iteration off the host: late iterations run a few stragglers and last ~270 us, about what the host needs to
        launch them."""
        dev = self.environment.device
        cuda = dev.type == "cuda"
        graph, use_graph = None, self.use_graph
        if use_graph is None:
            use_graph = cuda and not self.learned and self.check_every % 2 == 0 and self.max_iters >= 8 * self.check_every \
                and os.environ.get("PG_ILQR_GRAPH", "1") != "0"
        flags = torch.zeros(2, dtype=torch.int32, pin_memory=True) if cuda else None
        it, k, pending = 0, 0, None
        while it < self.max_iters:
            n = min(self.check_every, self.max_iters - it)
            if use_graph and it > 0 and n % 2 == 0:
                if graph is None:
                    graph = self._capture_two_iterations()
                for _ in range(n // 2):
                    graph.replay()
                for l, per in self._graph_launches:
                    l.launches += per * (n // 2)
            else:
                for _ in range(n):
                    self.iterate()
            it += n
            if not cuda:
                if int(self._n_active.item()) == 0:
                    break
                continue
            slot = k % 2
            k += 1
            flags[slot:slot + 1].copy_(self._n_active.reshape(1), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            if pending is not None:
                pending[0].synchronize()
                live = int(flags[pending[1]])
                if self._trace is not None:
                    self._trace.append((it - n, live, time.perf_counter()))  # count after iteration it - n, seen now
                if live == 0:
                    break
                if not self._tail and not self.learned and live <= self.tail_threshold:
                    if self.fused_tail and self._solve_tail(self.max_iters - it):
                        # the fused tail solver (pg_ilqr_solve) has been enqueued behind the iterations already in flight: it
                        # finishes every trajectory that is still running when it starts, in one launch
                        break
                    self._tail, graph = True, None  # the count only falls: switch the line search for good (new graph)
            pending = (ev, slot)
        self._graph = graph  # keep it alive until the results have been read

    def _solve_tail(self, remaining):
        """Finish the solve with the fused per-trajectory solver (lib.ilqr_solve): every trajectory on the current work
        list is iterated to its stopping test by a warp of its own -- in rounds of `tail_round` iterations (one launch
        each, all enqueued now: `remaining` iterations at most are left), because every launch deals the trajectories that
        are still running to the warps anew, and a short list gives each a warp scheduler of its own.  Results and
        decisions are bit-identical to continuing with iterate() (same device functions).  Returns False where the fused
        kernel does not apply."""
        if self.learned or self.finite_diff or self.dtype != torch.float64 or self.lin_mode not in (0, 1) or self.lib.m != 1:
            # (m = 2: the Car's active-set sweep rounds differently once inlined into the fused kernel -- iteration counts then
            # differ by rounding-level decisions; it keeps the launch-per-step tail, which is bit-identical by the tests)
            return False
        R = max(1, int(self.tail_round))
        rounds = max(1, -(-int(remaining) // R))
        for _ in range(rounds):
            cur = self._cur
            work = (self._idx[cur], self._cnt[cur])
            if self._trace is not None:  # tools/time_fused_tail.py: (event, live count) at the start of every round
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                self._round_trace.append((ev, self._cnt[cur].clone()))
            ok = self.lib.ilqr_solve(self._sel, self._x0, self._xbar, self._ubar, self._KK, self._kk, self._KKc, self._kkc, self._bw,
                                     self._fw, self._st, work, max_count=min(self.B, max(self.tail_threshold, 1)), S=self.substeps,
                                     dt=self.dt, integrator=self.integrator, reg_type=self.regType, lin_mode=self.lin_mode,
                                     ulim=self._ulim(), terminal_cost=float(self.environment.terminal_cost),
                                     last_tried=self._last_tried, round_iters=R, work_out=(self._idx[1 - cur], self._cnt[1 - cur]))
            if not ok:
                return False
            self._cur = 1 - cur
            self._n_active = self._cnt[self._cur]
            self._have_gains = True
        return True

    def _capture_two_iterations(self):
        """CUDA graph of iterate(); iterate() -- raw capture on a side stream (`with torch.cuda.graph` would first empty
        the caching allocator).  Everything iterate() touches is preallocated, its launch arguments are constants."""
        dev = self.environment.device
        libs = [self.lib]
        before = [l.launches for l in libs]
        cur = self._cur
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), lib.workspace_stream(torch.cuda.current_stream(dev)):
            graph.capture_begin(capture_error_mode="thread_local")  # other threads (NCCL watchdog, samplers) may call CUDA
            try:
                self.iterate()
                self.iterate()
            finally:
                graph.capture_end()
        torch.cuda.current_stream(dev).wait_stream(side)
        assert self._cur == cur
        self._graph_launches = [(l, l.launches - b) for l, b in zip(libs, before)]
        for l, b in zip(libs, before):  # the capture launched nothing
            l.launches = b
        return graph

    def run_optim(self, host_results=True):
        """Trajectory optimisation (ilqr.py:399-499).  Returns (xx, uu, cost) as host arrays, like the reference; with
        `host_results=False` the results stay on the device (`_xbar`, `_ubar`, `_st["cost"]`; the properties `xx`, `uu`,
        `cost` fetch them on demand) and None is returned -- a 16,384-trajectory batch otherwise copies 66 MB per call."""
        start_time = time.time()
        st = self._st
        # a new call re-arms every trajectory: `for _ in range(1, max_iters + 1)` starts over
        self._arm_state(self.reset_mu)
        self._sel = self._select_params()
        self._iterate_until_done()
        if self.final_backpass:
            self._final_backpass()
        self.iterations = int(st["iters"].max().item())
        self.lib.check_queues()  # a line-search worker that gave up waiting would have left stale candidate costs
        if not host_results and not self.printing and not self.saving:
            return None
        xx, uu = self.xx, self.uu
        self.agent.set_history(uu)
        if self.printing:
            stat = np.atleast_1d(self.status)
            print('Iterations: %d | Final Cost-to-Go: %.2f | Runtime: %.2f min. | %s'
                  % (self.iterations, float(np.mean(self.cost)), (time.time() - start_time) / 60,
                     ", ".join("%s: %d" % (lib.STATUS_NAMES[s], int((stat == s).sum())) for s in sorted(set(stat.tolist())))))
        if self.saving:
            self.save()
        return xx, uu, self.cost

    def run(self, x0):
        """Apply the stored time-varying controller from x0 (ilqr.py:348-351)."""
        self.environment.reset(x0)
        self._x0 = self.environment._to_dev(self.environment._x0_host, self.xDim)
        r = self._forward_candidates([1.0], self._KK, self._kk)
        self._xbar.copy_(r["X"][0])
        self._ubar.copy_(r["U"][0])
        self._st["cost"].copy_(r["cost"][0])

    def run_disk(self, x0):
        """Load a controller written by `save()` -- this package's or the reference's (ilqr.py:646-656) -- and apply it
        from x0 (ilqr.py:353-396).  Files under `path + 'data/'`: `K_.npy [T,m,n]`, `k.npy [T,m,1]`, `uu.npy [T,m]`,
        `xx.npy [T+1,n]`, `alpha.npy`, `time_info.p` (pickled list of the T+1 time stamps); one plant, or the same
        arrays with a leading batch axis.  A controller stored at another step size is re-sampled with a zero-order hold
        (`interp1d(kind='previous')` in the reference; its own version of this branch raises when the stored final time
        carries rounding, so the intended semantics are implemented: sample k of the new grid takes the stored entry
        floor(k dt_new / dt_old))."""
        d = (self.path or "") + 'data/'
        if not (self.path and os.path.isfile(d + 'K_.npy') and os.path.isfile(d + 'time_info.p')):
            self.environment.reset(x0)
            print("iLQR controller couldn't be loaded. Running initial trajectory.")
            self.init_trajectory()
            return
        with open(d + 'time_info.p', 'rb') as fh:
            tt = pickle.load(fh)
        dt_old = float(tt[1] - tt[0])
        KK, kk, uu, xx = (np.load(d + f) for f in ('K_.npy', 'k.npy', 'uu.npy', 'xx.npy'))
        alpha = np.load(d + 'alpha.npy')
        lead = (self.B,) if self.batched else ()
        T, n, m = self.steps, self.xDim, self.uDim
        KK = KK.reshape(lead + (-1, m, n))
        kk, uu, xx = kk.reshape(lead + (-1, m)), uu.reshape(lead + (-1, m)), xx.reshape(lead + (-1, n))
        ax = len(lead)
        T_old = KK.shape[ax]
        if self.dt == dt_old:
            src = np.minimum(np.arange(T), T_old - 1)        # forward_pass clamps j = min(len(KK) - 1, i) (ilqr.py:205)
        else:
            src = np.minimum(np.floor(np.arange(T) * self.dt / dt_old + 1e-9).astype(int), T_old - 1)
        KK, kk, uu = np.take(KK, src, axis=ax), np.take(kk, src, axis=ax), np.take(uu, src, axis=ax)
        xx = np.take(xx, np.append(src, min(src[-1] + 1, T_old)), axis=ax)
        self.environment.reset(x0)
        self._x0 = self.environment._to_dev(self.environment._x0_host, n)
        dev = lambda a: torch.as_tensor(np.ascontiguousarray(np.moveaxis(a.reshape((self.B,) + a.shape[ax:]), 0, -1)),
                                        dtype=self.dtype, device=self.environment.device)
        self._xbar.copy_(dev(xx))
        self._ubar.copy_(dev(uu))
        self._KK.copy_(dev(KK))
        self._kk.copy_(dev(kk))
        self._st["alpha_best"].copy_(torch.as_tensor(np.broadcast_to(np.asarray(alpha, dtype=float), (self.B,)).copy(),
                                                     dtype=self.dtype, device=self.environment.device))
        self._have_gains = True
        r = self.lib.ilqr_forward(self._x0, self._xbar, self._ubar, self._KK, self._kk, alpha_dev=self._st["alpha_best"],
                                  S=self.substeps, dt=self.dt, integrator=self.integrator, ulim=self._ulim(),
                                  terminal_cost=float(self.environment.terminal_cost), want_traj=True) \
            if not self.learned else self._forward_candidates([float(np.asarray(alpha).ravel()[0])], self._KK, self._kk)
        self._xbar.copy_(r["X"][0])
        self._ubar.copy_(r["U"][0])
        self._st["cost"].copy_(r["cost"][0])
        self.agent.set_history(self.uu)

    def save(self):
        """Controller on disk in the reference's format (ilqr.py:646-656)."""
        d = self.path + 'data/'
        np.save(d + 'K_', self.KK)
        np.save(d + 'k', self.kk)
        np.save(d + 'uu', self.uu)
        np.save(d + 'xx', self.xx)
        np.save(d + 'alpha', self.current_alpha)
        with open(d + 'time_info.p', 'wb') as fh:
            pickle.dump(list(np.arange(0, self.steps + 1) * self.dt), fh)
