"""ctypes binding of one compiled problem library (include/pygent_b200.h).

The Python layer passes torch CUDA tensors as raw device pointers + sizes + the current CUDA
stream; torch is plumbing (memory, streams), the arithmetic is in the library.  Every entry point
raises on a non-zero return code -- there is no fallback path.
"""
import ctypes
import functools
import os

import numpy as np
import torch

from . import build

PG_F64, PG_F32 = 0, 1
INTEG_RK4, INTEG_EULER = 0, 1
LIN_COST_FD = 4  # or-ed into lin_mode: cost expansion by central differences (iLQR(finite_diff=True))
MAX_ALPHAS = 16
ST_RUNNING, ST_CONV_FUN, ST_CONV_GRAD, ST_MU_MAX, ST_MAX_ITERS = range(5)
STATUS_NAMES = {0: "running", 1: "converged: small improvement", 2: "converged: small gradient",
                3: "diverged: no improvement", 4: "maximum iterations"}

_vp, _i32, _i64, _f64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
_dp = ctypes.POINTER(ctypes.c_double)


class ProblemInfo(ctypes.Structure):
    _fields_ = [("abi_version", _i32), ("n", _i32), ("m", _i32), ("n_obs", _i32), ("has_ab", _i32),
                ("has_cost_derivs", _i32), ("cost_uses_xnext", _i32), ("reserved", _i32),
                ("name", ctypes.c_char * 64), ("key", ctypes.c_char * 32)]


class SelectParams(ctypes.Structure):
    _fields_ = [("n_alpha", _i32), ("parallel", _i32), ("iteration", _i32), ("max_iters", _i32),
                ("alphas", _f64 * MAX_ALPHAS), ("zmin", _f64), ("tol_fun", _f64), ("tol_grad", _f64),
                ("mu_min", _f64), ("mu_max", _f64), ("mu_d0", _f64)]


class SolveArgs(ctypes.Structure):
    """pg_ilqr_solve_t (include/pygent_b200.h): every array of the fused tail solver."""
    _fields_ = [("dtype", _i32), ("integrator", _i32), ("T", _i32), ("S", _i32), ("n_alpha", _i32), ("reg_type", _i32),
                ("lin_mode", _i32), ("constrained", _i32), ("B", _i64), ("dt", _f64), ("terminal_cost", _f64), ("ulim", _f64 * 2),
                ("select", SelectParams), ("x0", _vp), ("xbar", _vp), ("ubar", _vp), ("KK", _vp), ("kk", _vp), ("KKc", _vp),
                ("kkc", _vp), ("dV", _vp), ("gnorm", _vp), ("bw_success", _vp), ("cost_cand", _vp), ("diverged", _vp),
                ("terminated", _vp), ("xN", _vp), ("cost", _vp), ("mu", _vp), ("mu_d", _vp), ("dcost", _vp),
                ("success_gradient", _vp), ("status", _vp), ("iters", _vp), ("alpha_best", _vp), ("accept", _vp), ("active", _vp),
                ("last_tried", _vp), ("index", _vp), ("count", _vp), ("max_count", _i32), ("workspace", _vp),
                ("workspace_bytes", _i64), ("round_iters", _i32), ("index_out", _vp), ("count_out", _vp)]


# name -> argtypes (restype is always int); mirrors include/pygent_b200.h
class AlphaWindow(ctypes.Structure):
    """pg_alpha_window (include/pygent_b200.h): which alphas one launch of a split serial line search evaluates."""
    _fields_ = [("phase", _i32), ("n_first", _i32), ("small_count", _i32), ("reserved", _i32), ("live_count", ctypes.c_void_p)]


SIGNATURES = {
    "pg_get_problem_info": [ctypes.POINTER(ProblemInfo)],
    "pg_rollout": [_i32, _i32, _i64, _i32, _i32, _f64, _f64, _vp, _vp, _vp, _vp, _vp, _vp, _f64, _i32, _vp],
    "pg_rollout_balanced": [_i32, _i32, _i64, _i32, _i32, _f64, _f64, _vp, _vp, _vp, _vp, _vp, _vp, _f64, _i32, _vp, _i64, _vp],
    "pg_rollout_random": [_i32, _i32, _i64, _i32, _i32, _f64, _f64, _vp, ctypes.c_uint64, _i32, _dp, _vp, _vp, _vp, _vp, _f64, _i32,
                          _vp, _i64, _vp],
    "pg_random_actions": [_i32, _i64, _i32, ctypes.c_uint64, _i32, _dp, _vp, _vp],
    "pg_ilqr_solve": [ctypes.POINTER(SolveArgs), _vp],
    "pg_ilqr_terminal_value": [_i32, _i64, _i32, _f64, _f64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "pg_ilqr_forward": [_i32, _i32, _i64, _i32, _i32, _f64, _i32, _dp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _dp, _f64,
                        _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "pg_ilqr_forward_balanced": [_i32, _i32, _i64, _i32, _i32, _f64, _i32, _dp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _dp, _f64,
                                 _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp],
    "pg_ilqr_commit_chunked": [_i32, _i32, _i64, _i32, _i32, _f64, _i32, _dp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _dp, _vp, _vp,
                               _vp, _vp, _vp, _vp],
    "pg_ilqr_linesearch_filter": [_i32, _i64, _i32, _i32, _dp, _f64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp],
    "pg_ilqr_backward": [_i32, _i64, _i32, _f64, _i32, _i32, _vp, _vp, _vp, _dp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                         _vp],
    "pg_ilqr_select": [_i32, _i64, ctypes.POINTER(SelectParams), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                       _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "pg_mpc_action": [_i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _dp, _vp, _vp],
    "pg_mpc_shift": [_i32, _i32, _i64, _i32, _i32, _f64, _f64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _f64, _vp],
    "pg_ilqr_commit": [_i32, _i32, _i64, _i32, _i32, _f64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _dp, _vp, _vp, _vp, _vp,
                       _vp],
    "pg_ilqr_backward_ext": [_i32, _i64, _i32, _f64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _dp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                             _vp],
    "pg_traj_cost": [_i32, _i64, _i32, _f64, _f64, _i32, _vp, _vp, _f64, _vp, _vp, _vp, _vp, _vp, _vp],
    "pg_ilqr_commit_copy": [_i32, _i64, _i32, _i32, _dp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "pg_observe": [_i32, _i64, _vp, _vp, _vp],
    "pg_eval": [_i32, _i64, _i32] + [_vp] * 16,
}


class PygentB200Error(RuntimeError):
    pass


def _dtype_code(dtype):
    if dtype == torch.float64:
        return PG_F64
    if dtype == torch.float32:
        return PG_F32
    raise TypeError("pygent_b200 kernels compute in float64 or float32, got %s" % dtype)


def _ptr(t, dtype=None, shape=None, name="tensor"):
    """Device pointer of a contiguous CUDA tensor (None -> NULL), with shape/dtype checks."""
    if t is None:
        return None
    if not t.is_cuda:
        raise PygentB200Error("%s must live on a CUDA device (there is no CPU path)" % name)
    if not t.is_contiguous():
        raise PygentB200Error("%s must be contiguous" % name)
    if dtype is not None and t.dtype != dtype:
        raise PygentB200Error("%s: expected dtype %s, got %s" % (name, dtype, t.dtype))
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise PygentB200Error("%s: expected shape %s, got %s" % (name, tuple(shape), tuple(t.shape)))
    return ctypes.c_void_p(t.data_ptr())


def _host_doubles(vals):
    if vals is None:
        return None
    vals = [float(v) for v in vals]
    return (ctypes.c_double * len(vals))(*vals)


_WS_STREAM_ALIAS = None


class workspace_stream(object):
    """While a CUDA graph is captured on a side stream that is forked from and joined back into `stream`, queue
    workspaces are looked up as if the calls were made on `stream` (the captured kernels logically run there), so the
    capture allocates nothing and the graph uses the buffers the eager iterations already own."""

    def __init__(self, stream):
        self.id = stream.cuda_stream

    def __enter__(self):
        global _WS_STREAM_ALIAS
        self.prev, _WS_STREAM_ALIAS = _WS_STREAM_ALIAS, self.id

    def __exit__(self, *exc):
        global _WS_STREAM_ALIAS
        _WS_STREAM_ALIAS = self.prev


class ProblemLibrary:
    """One loaded `pg_<system>_<key>.so`."""

    def __init__(self, path):
        self.path = path
        self.dll = ctypes.CDLL(path)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(self.dll, name)  # AttributeError if the library does not export the ABI
            fn.argtypes = argtypes
            fn.restype = ctypes.c_int
        self.dll.pg_rollout_workspace_bytes.argtypes = [_i64, _i32]
        self.dll.pg_rollout_workspace_bytes.restype = _i64
        self.dll.pg_ilqr_forward_workspace_bytes.argtypes = [_i64, _i32, _i32]
        self.dll.pg_ilqr_forward_workspace_bytes.restype = _i64
        self.dll.pg_ilqr_solve_capacity.argtypes = []
        self.dll.pg_ilqr_solve_capacity.restype = _i32
        self.dll.pg_ilqr_solve_workspace_bytes.argtypes = [_i32, _i32]
        self.dll.pg_ilqr_solve_workspace_bytes.restype = _i64
        self.dll.pg_ilqr_num_snapshots.argtypes = [_i32]
        self.dll.pg_ilqr_num_snapshots.restype = _i32
        self._workspace = {}
        self._retired = []
        info = ProblemInfo()
        self._check(self.dll.pg_get_problem_info(ctypes.byref(info)), "pg_get_problem_info")
        self.info = info
        self.n, self.m, self.n_obs = info.n, info.m, info.n_obs
        self.has_ab = bool(info.has_ab)
        self.name = info.name.decode()
        self.key = info.key.decode()
        self.launches = 0  # kernels launched through this handle (bench.py's gpu_launches)

    @staticmethod
    def _check(rc, what):
        if rc == 0:
            return
        if rc > 0:
            raise PygentB200Error("%s: CUDA error %d (%s)" % (what, rc, _cuda_error_name(rc)))
        raise PygentB200Error("%s: %s" % (what, {-1: "bad argument", -2: "not supported by this problem"}.get(rc, rc)))

    def _queue_workspace(self, kind, dev, nbytes):
        """Ticket-queue scratch of the balanced kernels: one buffer per (kernel kind, device, STREAM) -- this handle is
        shared by every solver and environment of the problem, and calls on different streams must not share counters.
        A buffer that has to grow is replaced, but the old one is kept alive: its address may be baked into a captured
        CUDA graph.  Bytes [0, 4) are the sticky status word the workers set when they give up waiting (queue_status)."""
        key = (kind, dev, _WS_STREAM_ALIAS if _WS_STREAM_ALIAS is not None else torch.cuda.current_stream(dev).cuda_stream)
        ws = self._workspace.get(key)
        if ws is None or ws.numel() < nbytes:
            if ws is not None:
                self._retired.append(ws)
            ws = self._workspace[key] = torch.zeros(max(nbytes, 1 << 16), dtype=torch.uint8, device=dev)
        return ws

    def queue_status_words(self):
        """int32 views of the status words of every live queue workspace (device tensors; reading them synchronises)."""
        return [ws[:4].view(torch.int32) for k, ws in self._workspace.items() if k[0] != "solve"] + \
               [ws[:4].view(torch.int32) for ws in self._retired]

    def check_queues(self):
        """Raise if a worker of a balanced (ticket-queue) kernel ever gave up waiting for its unit: the results of that
        launch are incomplete.  Never expected (GPU time-slicing under a debugger or MPS contention could do it); callers
        that already synchronise with the device check here so that the failure is loud instead of silent."""
        for w in self.queue_status_words():
            if int(w.item()) != 0:
                w.zero_()
                raise PygentB200Error("%s: a queue worker timed out waiting for its work unit; the results of that launch "
                                      "are incomplete" % self.name)

    @staticmethod
    def _work(work, B):
        if work is None:
            return None, None
        index, count = work
        return _ptr(index, torch.int32, (B,), "work index"), _ptr(count, torch.int32, (1,), "work count")

    @staticmethod
    def _stream():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    # -- K1 ------------------------------------------------------------------------------------
    def rollout(self, x0, U=None, T=None, S=4, dt=0.02, t0=0.0, integrator=INTEG_RK4, history=False,
                want_cost=True, terminal_cost=0.0, add_final_cost=False, accumulate_cost=False, out=None,
                balanced=None, actions=None):
        """x0 [n,B], U [T,m,B] or None -> dict(xN [n,B], cost [B], terminated [B] u8, X [T+1,n,B]).

        balanced: use the work-queue kernel (pg_rollout_balanced); default: when the batch fills the GPU but
        is not a whole number of warps per scheduler and the horizon has several chunks.
        actions = (seed, u_max [m][, k_offset]): the controls are drawn inside the kernel from the Philox stream of
        pygent_b200/random_actions.py (pg_rollout_random) instead of being read from U."""
        n, m = self.n, self.m
        dtype, dev = x0.dtype, x0.device
        B = x0.shape[1]
        if U is not None:
            T = U.shape[0]
        if T is None:
            raise PygentB200Error("rollout needs U or T")
        out = out or {}
        xN = out.get("xN")
        if xN is None:
            xN = torch.empty((n, B), dtype=dtype, device=dev)
        cost = out.get("cost")
        if cost is None and want_cost:
            cost = torch.empty((B,), dtype=dtype, device=dev)
        term = out.get("terminated")
        if term is None:
            term = torch.empty((B,), dtype=torch.uint8, device=dev)
        X = out.get("X")
        if X is None and history:
            X = torch.empty((T + 1, n, B), dtype=dtype, device=dev)
        flags = int(add_final_cost) | (2 if accumulate_cost else 0)
        if balanced is None:
            balanced = use_balanced_rollout(B, T, dev)
        if actions is not None:
            if U is not None:
                raise PygentB200Error("rollout: pass U or actions=(seed, u_max), not both")
            seed, u_max = int(actions[0]), [float(v) for v in np.ravel(actions[1])]
            k_off = int(actions[2]) if len(actions) > 2 else 0
            if len(u_max) != m:
                raise PygentB200Error("actions: u_max needs %d entries" % m)
            ws, nbytes = None, 0
            if balanced and cost is not None and T >= 1 and B > 0:
                nbytes = int(self.dll.pg_rollout_workspace_bytes(B, T))
                ws = self._queue_workspace("ro", dev, nbytes)
            rc = self.dll.pg_rollout_random(_dtype_code(dtype), integrator, B, T, S, dt, t0, _ptr(x0, dtype, (n, B), "x0"),
                                            ctypes.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), k_off, _host_doubles(u_max),
                                            _ptr(X, dtype, (T + 1, n, B), "X"), _ptr(xN, dtype, (n, B), "xN"),
                                            _ptr(cost, dtype, (B,), "cost"), _ptr(term, torch.uint8, (B,), "terminated"),
                                            terminal_cost, flags, _ptr(ws) if ws is not None else None, nbytes, self._stream())
            self._check(rc, "pg_rollout_random")
            self.launches += 1
            return {"xN": xN, "cost": cost, "terminated": term, "X": X}
        if balanced and cost is not None and T >= 1 and B > 0:
            nbytes = int(self.dll.pg_rollout_workspace_bytes(B, T))
            ws = self._queue_workspace("ro", dev, nbytes)
            rc = self.dll.pg_rollout_balanced(_dtype_code(dtype), integrator, B, T, S, dt, t0,
                                              _ptr(x0, dtype, (n, B), "x0"), _ptr(U, dtype, (T, m, B), "U"),
                                              _ptr(X, dtype, (T + 1, n, B), "X"), _ptr(xN, dtype, (n, B), "xN"),
                                              _ptr(cost, dtype, (B,), "cost"), _ptr(term, torch.uint8, (B,), "terminated"),
                                              terminal_cost, flags, _ptr(ws), nbytes, self._stream())
            self._check(rc, "pg_rollout_balanced")
            self.launches += 1
            return {"xN": xN, "cost": cost, "terminated": term, "X": X}
        rc = self.dll.pg_rollout(_dtype_code(dtype), integrator, B, T, S, dt, t0,
                                 _ptr(x0, dtype, (n, B), "x0"), _ptr(U, dtype, (T, m, B), "U"),
                                 _ptr(X, dtype, (T + 1, n, B), "X"), _ptr(xN, dtype, (n, B), "xN"),
                                 _ptr(cost, dtype, (B,), "cost"), _ptr(term, torch.uint8, (B,), "terminated"),
                                 terminal_cost, flags, self._stream())
        self._check(rc, "pg_rollout")
        self.launches += 1
        return {"xN": xN, "cost": cost, "terminated": term, "X": X}

    def random_actions(self, B, T, seed, u_max, k_offset=0, dtype=torch.float64, device=None):
        """U [T, m, B] of the in-kernel action stream (pg_random_actions), materialised."""
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        U = torch.empty((T, self.m, B), dtype=dtype, device=dev)
        rc = self.dll.pg_random_actions(_dtype_code(dtype), B, T, ctypes.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), int(k_offset),
                                        _host_doubles([float(v) for v in np.ravel(u_max)]), _ptr(U, dtype, (T, self.m, B), "U"), self._stream())
        self._check(rc, "pg_random_actions")
        self.launches += 1
        return U

    # -- K2 ------------------------------------------------------------------------------------
    def ilqr_forward(self, x0, xbar, ubar, KK, kk, alphas=None, alpha_dev=None, S=4, dt=0.02,
                     integrator=INTEG_RK4, ulim=None, terminal_cost=0.0, want_traj=False, active=None, out=None,
                     work=None, balanced=None, window=None):
        """work: (index i32 [B], count i32 [1]) compacted list of trajectories to process, or None (all).
        window: (phase, n_first, small_count, live_count i32 [1] device) of a split serial line search, or None.
        balanced: use the work-queue kernel (pg_ilqr_forward_balanced; costs only); default: when there are more
        (32-trajectory group, alpha) pairs than the 3 x 4 x SM persistent workers."""
        n, m = self.n, self.m
        dtype, dev = x0.dtype, x0.device
        B = x0.shape[1]
        T = ubar.shape[0]
        Tk = KK.shape[0]
        n_alpha = 1 if alpha_dev is not None else len(alphas)
        out = out or {}
        cost = out.get("cost")
        if cost is None:
            cost = torch.empty((n_alpha, B), dtype=dtype, device=dev)
        div = out.get("diverged")
        if div is None:
            div = torch.empty((n_alpha, B), dtype=torch.uint8, device=dev)
        term = out.get("terminated")
        if term is None:
            term = torch.empty((n_alpha, B), dtype=torch.uint8, device=dev)
        X = out.get("X")
        Uo = out.get("U")
        xN = out.get("xN")  # [n_alpha, n, B]: terminal states of the candidates (MPC plan shift)
        xS = out.get("xS")  # [n_alpha, num_snapshots(T), n, B]: candidates at the chunk boundaries (chunk-parallel commit)
        xS_ptr = _ptr(xS, dtype, (n_alpha, self.num_snapshots(T), n, B), "xS") if xS is not None else None
        win = None
        if window is not None:
            win_struct = AlphaWindow(int(window[0]), int(window[1]), int(window[2]), 0, _ptr(window[3], torch.int32, (1,), "live_count"))
            win = ctypes.addressof(win_struct)  # read on the host during the call; win_struct lives until we return
        if want_traj and X is None:
            X = torch.empty((n_alpha, T + 1, n, B), dtype=dtype, device=dev)
            Uo = torch.empty((n_alpha, T, m, B), dtype=dtype, device=dev)
        if balanced is None:
            balanced = use_balanced_forward(B, n_alpha, T, dev)
        if balanced and not want_traj and X is None and active is None and T >= 1:
            if xN is None:
                xN = torch.empty((n_alpha, n, B), dtype=dtype, device=dev)
            nbytes = int(self.dll.pg_ilqr_forward_workspace_bytes(B, T, n_alpha))
            ws = self._queue_workspace("fw", dev, nbytes)
            rc = self.dll.pg_ilqr_forward_balanced(
                _dtype_code(dtype), integrator, B, T, S, dt, n_alpha, _host_doubles(alphas),
                _ptr(alpha_dev, dtype, (B,), "alpha_dev"), _ptr(x0, dtype, (n, B), "x0"),
                _ptr(xbar, dtype, (T + 1, n, B), "xbar"), _ptr(ubar, dtype, (T, m, B), "ubar"),
                _ptr(KK, dtype, (Tk, m, n, B), "KK"), _ptr(kk, dtype, (Tk, m, B), "kk"), Tk, _host_doubles(ulim),
                terminal_cost, _ptr(cost, dtype, (n_alpha, B), "cost_out"), _ptr(div, torch.uint8, (n_alpha, B), "diverged_out"),
                _ptr(term, torch.uint8, (n_alpha, B), "terminated_out"), _ptr(xN, dtype, (n_alpha, n, B), "xN_out"), xS_ptr,
                *self._work(work, B), _ptr(ws), nbytes, win, self._stream())
            self._check(rc, "pg_ilqr_forward_balanced")
            self.launches += 1
            return {"cost": cost, "diverged": div, "terminated": term, "X": None, "U": None, "xN": xN}
        rc = self.dll.pg_ilqr_forward(
            _dtype_code(dtype), integrator, B, T, S, dt, n_alpha, _host_doubles(alphas),
            _ptr(alpha_dev, dtype, (B,), "alpha_dev"), _ptr(x0, dtype, (n, B), "x0"),
            _ptr(xbar, dtype, (T + 1, n, B), "xbar"), _ptr(ubar, dtype, (T, m, B), "ubar"),
            _ptr(KK, dtype, (Tk, m, n, B), "KK"), _ptr(kk, dtype, (Tk, m, B), "kk"), Tk, _host_doubles(ulim),
            terminal_cost, _ptr(X, dtype, (n_alpha, T + 1, n, B), "X_out"), _ptr(Uo, dtype, (n_alpha, T, m, B), "U_out"),
            _ptr(cost, dtype, (n_alpha, B), "cost_out"), _ptr(div, torch.uint8, (n_alpha, B), "diverged_out"),
            _ptr(term, torch.uint8, (n_alpha, B), "terminated_out"), _ptr(xN, dtype, (n_alpha, n, B), "xN_out"), xS_ptr,
            _ptr(active, torch.uint8, (B,), "active"), *self._work(work, B), win, self._stream())
        self._check(rc, "pg_ilqr_forward")
        self.launches += 1
        return {"cost": cost, "diverged": div, "terminated": term, "X": X, "U": Uo, "xN": xN}

    def terminal_value(self, xbar, dt, t_final, lin_mode=0, work=None):
        """pg_ilqr_terminal_value: xbar [T+1, n, B] -> dict(P [n, n, B], x_eq [n, B], ok [B] u8); None if this problem /
        dtype is not supported on the device (the caller then uses the host route)."""
        dtype, dev = xbar.dtype, xbar.device
        T, n, B = xbar.shape[0] - 1, self.n, xbar.shape[2]
        P = torch.empty((n, n, B), dtype=dtype, device=dev)
        xe = torch.empty((n, B), dtype=dtype, device=dev)
        ok = torch.zeros((B,), dtype=torch.uint8, device=dev)
        rc = self.dll.pg_ilqr_terminal_value(_dtype_code(dtype), B, T, dt, t_final, lin_mode & 3, _ptr(xbar, dtype, (T + 1, n, B), "xbar"),
                                             _ptr(P, dtype, (n, n, B), "P"), _ptr(xe, dtype, (n, B), "x_eq"), _ptr(ok, torch.uint8, (B,), "ok"),
                                             *self._work(work, B), self._stream())
        if rc == -2:
            return None
        self._check(rc, "pg_ilqr_terminal_value")
        self.launches += 1
        return {"P": P, "x_eq": xe, "ok": ok}

    # -- fused tail solver -----------------------------------------------------------------------
    def solve_capacity(self):
        """Trajectories pg_ilqr_solve works on at once (resident warps of the fused solver on this device)."""
        if not torch.cuda.is_available():
            return 0
        return int(self.dll.pg_ilqr_solve_capacity())

    def ilqr_solve(self, sel, x0, xbar, ubar, KK, kk, KKc, kkc, bw, fw, st, work, max_count, S=4, dt=0.02, integrator=INTEG_RK4,
                   reg_type=2, lin_mode=0, ulim=None, terminal_cost=0.0, last_tried=None, round_iters=0, work_out=None):
        """pg_ilqr_solve: ONE launch that finishes every trajectory of the work list `work` = (index, count) -- a warp per
        trajectory loops backward sweep / 11 candidates side by side / decision / commit until the trajectory stops -- or,
        with `round_iters` > 0, gives each at most that many iterations and leaves the list of those still running in
        `work_out` = (index, count) for the next call.
        bw = dict(dV, gnorm, success), fw = dict(cost, diverged, terminated, xN), st = the solver state dict of ilqr_select.
        Returns False if this configuration is not supported by the fused kernel (the caller keeps iterating)."""
        dtype, dev = x0.dtype, x0.device
        T, B = ubar.shape[0], x0.shape[1]
        nbytes = int(self.dll.pg_ilqr_solve_workspace_bytes(T, int(max_count)))
        key = ("solve", dev, _WS_STREAM_ALIAS if _WS_STREAM_ALIAS is not None else torch.cuda.current_stream(dev).cuda_stream)
        ws = self._workspace.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = self._workspace[key] = torch.zeros(max(nbytes, 1 << 16), dtype=torch.uint8, device=dev)
        a = SolveArgs()
        a.dtype, a.integrator, a.T, a.S, a.n_alpha = _dtype_code(dtype), integrator, T, S, sel.n_alpha
        a.reg_type, a.lin_mode, a.constrained, a.B = reg_type, lin_mode, 1 if ulim is not None else 0, B
        a.dt, a.terminal_cost = dt, terminal_cost
        for i, v in enumerate(ulim or []):
            a.ulim[i] = float(v)
        a.select = sel
        p = lambda t: None if t is None else t.data_ptr()
        a.x0, a.xbar, a.ubar, a.KK, a.kk, a.KKc, a.kkc = p(x0), p(xbar), p(ubar), p(KK), p(kk), p(KKc), p(kkc)
        a.dV, a.gnorm, a.bw_success = p(bw["dV"]), p(bw["gnorm"]), p(bw["success"])
        a.cost_cand, a.diverged, a.terminated, a.xN = p(fw["cost"]), p(fw["diverged"]), p(fw.get("terminated")), p(fw.get("xN"))
        a.cost, a.mu, a.mu_d, a.dcost = p(st["cost"]), p(st["mu"]), p(st["mu_d"]), p(st["dcost"])
        a.success_gradient, a.status, a.iters = p(st["success_gradient"]), p(st["status"]), p(st["iters"])
        a.alpha_best, a.accept, a.active, a.last_tried = p(st["alpha_best"]), p(st["accept"]), p(st["active"]), p(last_tried)
        a.index, a.count, a.max_count = p(work[0]), p(work[1]), int(max_count)
        a.workspace, a.workspace_bytes = ws.data_ptr(), nbytes
        if round_iters > 0:
            a.round_iters, a.index_out, a.count_out = int(round_iters), p(work_out[0]), p(work_out[1])
        rc = self.dll.pg_ilqr_solve(ctypes.byref(a), self._stream())
        if rc == -2:
            return False
        self._check(rc, "pg_ilqr_solve")
        self.launches += 1
        return True

    # -- K3 ------------------------------------------------------------------------------------
    def ilqr_backward(self, xbar, ubar, mu, dt=0.02, reg_type=2, lin_mode=None, ulim=None, active=None, out=None,
                      work=None, Vxx_T=None):
        n, m = self.n, self.m
        dtype, dev = xbar.dtype, xbar.device
        T, B = ubar.shape[0], ubar.shape[2]
        if lin_mode is None:
            lin_mode = 0 if self.has_ab else 1
        out = out or {}
        KK = out.get("KK")
        if KK is None:
            KK = torch.zeros((T, m, n, B), dtype=dtype, device=dev)
        kk = out.get("kk")
        if kk is None:
            kk = torch.zeros((T, m, B), dtype=dtype, device=dev)
        dV = out.get("dV")
        if dV is None:
            dV = torch.empty((2, B), dtype=dtype, device=dev)
        gnorm = out.get("gnorm")
        if gnorm is None:
            gnorm = torch.empty((B,), dtype=dtype, device=dev)
        success = out.get("success")
        if success is None:
            success = torch.empty((B,), dtype=torch.uint8, device=dev)
        rc = self.dll.pg_ilqr_backward(
            _dtype_code(dtype), B, T, dt, reg_type, lin_mode, _ptr(mu, dtype, (B,), "mu"),
            _ptr(xbar, dtype, (T + 1, n, B), "xbar"), _ptr(ubar, dtype, (T, m, B), "ubar"), _host_doubles(ulim),
            _ptr(Vxx_T, dtype, (n, n, B), "Vxx_T"),
            _ptr(KK, dtype, (T, m, n, B), "KK"), _ptr(kk, dtype, (T, m, B), "kk"), _ptr(dV, dtype, (2, B), "dV"),
            _ptr(gnorm, dtype, (B,), "gnorm"), _ptr(success, torch.uint8, (B,), "success"),
            _ptr(active, torch.uint8, (B,), "active"), *self._work(work, B), self._stream())
        self._check(rc, "pg_ilqr_backward")
        self.launches += 1
        return {"KK": KK, "kk": kk, "dV": dV, "gnorm": gnorm, "success": success}

    # -- K4 ------------------------------------------------------------------------------------
    def ilqr_select(self, params, cost_cand, diverged, dV, gnorm, bw_success, state, n_active=None, work=None,
                    index_out=None, last_tried=None):
        """state: dict of per-trajectory tensors cost, mu, mu_d, dcost, success_gradient, status, iters,
        alpha_best, accept, active (updated in place)."""
        dtype = cost_cand.dtype
        B = cost_cand.shape[1]
        u8, i32 = torch.uint8, torch.int32
        rc = self.dll.pg_ilqr_select(
            _dtype_code(dtype), B, ctypes.byref(params), _ptr(cost_cand, dtype, None, "cost_cand"),
            _ptr(diverged, u8, tuple(cost_cand.shape), "diverged"), _ptr(dV, dtype, (2, B), "dV"),
            _ptr(gnorm, dtype, (B,), "gnorm"), _ptr(bw_success, u8, (B,), "bw_success"),
            _ptr(state["cost"], dtype, (B,), "cost"), _ptr(state["mu"], dtype, (B,), "mu"),
            _ptr(state["mu_d"], dtype, (B,), "mu_d"), _ptr(state["dcost"], dtype, (B,), "dcost"),
            _ptr(state["success_gradient"], u8, (B,), "success_gradient"), _ptr(state["status"], i32, (B,), "status"),
            _ptr(state["iters"], i32, (B,), "iters"), _ptr(state["alpha_best"], dtype, (B,), "alpha_best"),
            _ptr(state["accept"], u8, (B,), "accept"), _ptr(state["active"], u8, (B,), "active"),
            _ptr(n_active, i32, None, "n_active"), *self._work(work, B), _ptr(index_out, i32, (B,), "index_out"),
            _ptr(last_tried, i32, (B,), "last_tried"), self._stream())
        self._check(rc, "pg_ilqr_select")
        self.launches += 1

    def ilqr_commit(self, x0, xbar, ubar, KK, kk, KK_acc, kk_acc, alpha_best, accept, S=4, dt=0.02,
                    integrator=INTEG_RK4, ulim=None, work=None):
        n, m = self.n, self.m
        dtype = x0.dtype
        B = x0.shape[1]
        T = ubar.shape[0]
        rc = self.dll.pg_ilqr_commit(
            _dtype_code(dtype), integrator, B, T, S, dt, _ptr(alpha_best, dtype, (B,), "alpha_best"),
            _ptr(accept, torch.uint8, (B,), "accept"), _ptr(x0, dtype, (n, B), "x0"),
            _ptr(xbar, dtype, (T + 1, n, B), "xbar"), _ptr(ubar, dtype, (T, m, B), "ubar"),
            _ptr(KK, dtype, (T, m, n, B), "KK"), _ptr(kk, dtype, (T, m, B), "kk"), _host_doubles(ulim),
            _ptr(KK_acc, dtype, (T, m, n, B), "KK_acc"), _ptr(kk_acc, dtype, (T, m, B), "kk_acc"), *self._work(work, B),
            self._stream())
        self._check(rc, "pg_ilqr_commit")
        self.launches += 1

    # -- dynamics integrated / linearised outside the library (learned MLP) --------------------------
    def ilqr_backward_ext(self, xbar, ubar, mu, A, Bm, dt=0.02, reg_type=2, ulim=None, out=None, work=None, Vxx_T=None,
                          cost_fd=False):
        """Riccati sweep with caller-supplied continuous-time Jacobians A [T,n,n,B], Bm [T,n,m,B]."""
        n, m = self.n, self.m
        dtype, dev = xbar.dtype, xbar.device
        T, B = ubar.shape[0], ubar.shape[2]
        out = out or {}
        g = lambda k, shape, dt_=dtype, zero=False: out[k] if out.get(k) is not None else (
            torch.zeros(shape, dtype=dt_, device=dev) if zero else torch.empty(shape, dtype=dt_, device=dev))
        KK, kk = g("KK", (T, m, n, B), zero=True), g("kk", (T, m, B), zero=True)
        dV, gnorm, success = g("dV", (2, B)), g("gnorm", (B,)), g("success", (B,), torch.uint8)
        rc = self.dll.pg_ilqr_backward_ext(
            _dtype_code(dtype), B, T, dt, reg_type, LIN_COST_FD if cost_fd else 0, _ptr(mu, dtype, (B,), "mu"),
            _ptr(xbar, dtype, (T + 1, n, B), "xbar"),
            _ptr(ubar, dtype, (T, m, B), "ubar"), _ptr(A, dtype, (T, n, n, B), "A"), _ptr(Bm, dtype, (T, n, m, B), "Bm"),
            _host_doubles(ulim), _ptr(Vxx_T, dtype, (n, n, B), "Vxx_T"), _ptr(KK, dtype, (T, m, n, B), "KK"),
            _ptr(kk, dtype, (T, m, B), "kk"), _ptr(dV, dtype, (2, B), "dV"), _ptr(gnorm, dtype, (B,), "gnorm"),
            _ptr(success, torch.uint8, (B,), "success"), *self._work(work, B), self._stream())
        self._check(rc, "pg_ilqr_backward_ext")
        self.launches += 1
        return {"KK": KK, "kk": kk, "dV": dV, "gnorm": gnorm, "success": success}

    def traj_cost(self, X, U, dt=0.02, t0=0.0, terminal_cost=0.0, out=None, work=None):
        """X [n_alpha,T+1,n,B], U [n_alpha,T,m,B] -> dict(cost, diverged, terminated), each [n_alpha,B]."""
        n, m = self.n, self.m
        dtype, dev = X.dtype, X.device
        A, T, B = X.shape[0], X.shape[1] - 1, X.shape[3]
        out = out or {}
        cost = out.get("cost") if out.get("cost") is not None else torch.empty((A, B), dtype=dtype, device=dev)
        div = out.get("diverged") if out.get("diverged") is not None else torch.empty((A, B), dtype=torch.uint8, device=dev)
        term = out.get("terminated") if out.get("terminated") is not None else torch.empty((A, B), dtype=torch.uint8, device=dev)
        rc = self.dll.pg_traj_cost(_dtype_code(dtype), B, T, dt, t0, A, _ptr(X, dtype, (A, T + 1, n, B), "X"),
                                   _ptr(U, dtype, (A, T, m, B), "U"), terminal_cost, _ptr(cost, dtype, (A, B), "cost_out"),
                                   _ptr(div, torch.uint8, (A, B), "diverged_out"), _ptr(term, torch.uint8, (A, B), "terminated_out"),
                                   *self._work(work, B), self._stream())
        self._check(rc, "pg_traj_cost")
        self.launches += 1
        return {"cost": cost, "diverged": div, "terminated": term}

    def ilqr_commit_copy(self, alphas, alpha_best, accept, X_cand, U_cand, xbar, ubar, KK, kk, KK_acc, kk_acc, work=None):
        n, m = self.n, self.m
        dtype = xbar.dtype
        A, T, B = X_cand.shape[0], ubar.shape[0], ubar.shape[2]
        rc = self.dll.pg_ilqr_commit_copy(
            _dtype_code(dtype), B, T, A, _host_doubles(alphas), _ptr(alpha_best, dtype, (B,), "alpha_best"),
            _ptr(accept, torch.uint8, (B,), "accept"), _ptr(X_cand, dtype, (A, T + 1, n, B), "X_cand"),
            _ptr(U_cand, dtype, (A, T, m, B), "U_cand"), _ptr(xbar, dtype, (T + 1, n, B), "xbar"),
            _ptr(ubar, dtype, (T, m, B), "ubar"), _ptr(KK, dtype, (T, m, n, B), "KK"), _ptr(kk, dtype, (T, m, B), "kk"),
            _ptr(KK_acc, dtype, (T, m, n, B), "KK_acc"), _ptr(kk_acc, dtype, (T, m, B), "kk_acc"), *self._work(work, B),
            self._stream())
        self._check(rc, "pg_ilqr_commit_copy")
        self.launches += 1

    # -- receding-horizon control ------------------------------------------------------------------
    def mpc_action(self, KK, kk, xbar, ubar, alpha, x, ulim, noise=None, out=None):
        """u [m,B] = clip(K_0 (x - xbar_0) + ubar_0 + alpha k_0 + noise, +-ulim) (nmpc.py:197-206)."""
        n, m = self.n, self.m
        dtype = x.dtype
        B = x.shape[1]
        u = out if out is not None else torch.empty((m, B), dtype=dtype, device=x.device)
        rc = self.dll.pg_mpc_action(_dtype_code(dtype), B, _ptr(KK, dtype, None, "KK"), _ptr(kk, dtype, None, "kk"),
                                    _ptr(xbar, dtype, None, "xbar"), _ptr(ubar, dtype, None, "ubar"),
                                    _ptr(alpha, dtype, (B,), "alpha"), _ptr(x, dtype, (n, B), "x"),
                                    _ptr(noise, dtype, (m, B), "noise"), _host_doubles(ulim), _ptr(u, dtype, (m, B), "u_out"),
                                    self._stream())
        self._check(rc, "pg_mpc_action")
        self.launches += 1
        return u

    def mpc_shift(self, xbar, ubar, KK, kk, cost, xe, dt=0.02, S=4, integrator=INTEG_RK4, t_env=0.0, xN_cand=None,
                  last_tried=None, x_new_ext=None, terminal_cost=0.0):
        """shift_planner (nmpc.py:257-277) in place; see include/pygent_b200.h."""
        n, m = self.n, self.m
        dtype = xbar.dtype
        T, B = ubar.shape[0], ubar.shape[2]
        n_alpha = xN_cand.shape[0] if xN_cand is not None else 0
        rc = self.dll.pg_mpc_shift(_dtype_code(dtype), integrator, B, T, S, dt, t_env, _ptr(xbar, dtype, (T + 1, n, B), "xbar"),
                                   _ptr(ubar, dtype, (T, m, B), "ubar"), _ptr(KK, dtype, (T, m, n, B), "KK"),
                                   _ptr(kk, dtype, (T, m, B), "kk"), _ptr(cost, dtype, (B,), "cost"), _ptr(xe, dtype, (n, B), "xe"),
                                   _ptr(xN_cand, dtype, None, "xN_cand"), n_alpha, _ptr(last_tried, torch.int32, (B,), "last_tried"),
                                   _ptr(x_new_ext, dtype, (n, B), "x_new_ext"), terminal_cost, self._stream())
        self._check(rc, "pg_mpc_shift")
        self.launches += 1

    def linesearch_filter(self, phase, bw_success, index_out, count_out, work=None, n_first=0, alphas=None, zmin=0.0,
                          cost=None, dV=None, cost_cand=None, diverged=None, small_count=0):
        """Compacted work list of one phase of the two-launch serial line search (see include/pygent_b200.h)."""
        B = bw_success.shape[0]
        dtype = cost.dtype if cost is not None else torch.float64
        rc = self.dll.pg_ilqr_linesearch_filter(
            _dtype_code(dtype), B, phase, n_first, _host_doubles(alphas), zmin, _ptr(cost, dtype, (B,), "cost"),
            _ptr(dV, dtype, (2, B), "dV"), _ptr(bw_success, torch.uint8, (B,), "bw_success"), _ptr(cost_cand, dtype, None, "cost_cand"),
            _ptr(diverged, torch.uint8, None, "diverged"), *self._work(work, B), _ptr(index_out, torch.int32, (B,), "index_out"),
            _ptr(count_out, torch.int32, (1,), "count_out"), int(small_count), self._stream())
        self._check(rc, "pg_ilqr_linesearch_filter")
        self.launches += 1

    def num_snapshots(self, T):
        return int(self.dll.pg_ilqr_num_snapshots(int(T)))

    def ilqr_commit_chunked(self, alphas, x0, xbar, ubar, KK, kk, KK_acc, kk_acc, alpha_best, accept, xS,
                            S=4, dt=0.02, integrator=INTEG_RK4, ulim=None, work=None):
        """pg_ilqr_commit with the winner's time chunks integrated in parallel from the snapshots xS."""
        n, m = self.n, self.m
        dtype = x0.dtype
        B, T, A = x0.shape[1], ubar.shape[0], len(alphas)
        rc = self.dll.pg_ilqr_commit_chunked(
            _dtype_code(dtype), integrator, B, T, S, dt, A, _host_doubles(alphas), _ptr(alpha_best, dtype, (B,), "alpha_best"),
            _ptr(accept, torch.uint8, (B,), "accept"), _ptr(x0, dtype, (n, B), "x0"), _ptr(xbar, dtype, (T + 1, n, B), "xbar"),
            _ptr(ubar, dtype, (T, m, B), "ubar"), _ptr(KK, dtype, (T, m, n, B), "KK"), _ptr(kk, dtype, (T, m, B), "kk"),
            _host_doubles(ulim), _ptr(xS, dtype, (A, self.num_snapshots(T), n, B), "xS"),
            _ptr(KK_acc, dtype, (T, m, n, B), "KK_acc"), _ptr(kk_acc, dtype, (T, m, B), "kk_acc"), *self._work(work, B),
            self._stream())
        self._check(rc, "pg_ilqr_commit_chunked")
        self.launches += 1

    def observe(self, x):
        B = x.shape[1]
        o = torch.empty((self.n_obs, B), dtype=x.dtype, device=x.device)
        self._check(self.dll.pg_observe(_dtype_code(x.dtype), B, _ptr(x, None, (self.n, B), "x"), _ptr(o), self._stream()),
                    "pg_observe")
        self.launches += 1
        return o

    def eval(self, x, u, t=None, lin_mode=None, what=("f", "A", "B", "c", "cx", "cu", "Cxx", "Cuu", "Cxu", "cf", "cfx", "Cfxx")):
        """Point evaluation of the generated functions; returns dict name -> tensor [..., B]."""
        n, m = self.n, self.m
        dtype, dev = x.dtype, x.device
        B = x.shape[1]
        if lin_mode is None:
            lin_mode = 0 if self.has_ab else 1
        shapes = {"f": (n, B), "A": (n, n, B), "B": (n, m, B), "c": (B,), "cx": (n, B), "cu": (m, B), "Cxx": (n, n, B),
                  "Cuu": (m, m, B), "Cxu": (n, m, B), "cf": (B,), "cfx": (n, B), "Cfxx": (n, n, B)}
        if not self.info.has_cost_derivs:
            what = [w for w in what if w not in ("cx", "cu", "Cxx", "Cuu", "Cxu")]
        outs = {k: torch.empty(shapes[k], dtype=dtype, device=dev) for k in what}
        order = ["f", "A", "B", "c", "cx", "cu", "Cxx", "Cuu", "Cxu", "cf", "cfx", "Cfxx"]
        rc = self.dll.pg_eval(_dtype_code(dtype), B, lin_mode, _ptr(x, dtype, (n, B), "x"), _ptr(u, dtype, (m, B), "u"),
                              _ptr(t, dtype, (B,), "t"), *[_ptr(outs.get(k)) for k in order], self._stream())
        self._check(rc, "pg_eval")
        self.launches += 1
        return outs


@functools.lru_cache(maxsize=None)
def _sm_count(index):
    return torch.cuda.get_device_properties(index).multi_processor_count


def use_balanced_rollout(B, T, dev):
    """Work-queue kernel pays off when the GPU is full (>= 2 warps per scheduler) but the warps do not divide
    evenly over the 4 x SM schedulers, and the horizon spans several 25-step chunks."""
    if dev.type != "cuda":
        return False
    sched = 4 * _sm_count(dev.index if dev.index is not None else torch.cuda.current_device())
    warps = (B + 31) // 32
    if warps < 2 * sched or T < 100:
        return False
    per = warps / sched
    return (-(-warps // sched)) / per > 1.04


def use_balanced_forward(B, n_alpha, T, dev):
    """Queue kernel for the line search when its (group, alpha) pairs outnumber the persistent workers (3 per warp
    scheduler) and the horizon spans several 20-step chunks."""
    if dev.type != "cuda" or T < 40:
        return False
    workers = 3 * 4 * _sm_count(dev.index if dev.index is not None else torch.cuda.current_device())
    return ((B + 31) // 32) * n_alpha > workers


def split_search_threshold(dev):
    """Live trajectories below which a serial line search is evaluated in one launch (pg_alpha_window.small_count):
    two dependent rollout chains only pay when the first one fills the GPU -- about 64 trajectories per SM."""
    v = os.environ.get("PG_LS_SMALL")
    if v is not None:
        return int(v)
    if dev.type != "cuda":
        return 1 << 30
    return 64 * _sm_count(dev.index if dev.index is not None else torch.cuda.current_device())


def _cuda_error_name(rc):
    try:
        cudart = torch.cuda.cudart()
        return cudart.cudaGetErrorString(rc) if hasattr(cudart, "cudaGetErrorString") else "cudaError"
    except Exception:
        return "cudaError"


@functools.lru_cache(maxsize=None)
def _load(path):
    return ProblemLibrary(path)


def load_problem(spec):
    """Build (if needed) and load the library of `spec`; raises if neither is possible."""
    return _load(build.build_problem(spec))
