"""GPU path of the learned-MLP dynamics (include/pygent_b200_mlp.h): library build, ctypes binding, and the
environment that integrates `MBRL.ode` (pygent/algorithms/mbrl.py:125-130) with forward Euler.

`precision="fp32"`: layer 2 in fp32 FFMA -- tracks the reference's fp32 torch network to ~1e-6;
`precision="f16x3"`: layer 2 on the tcgen05 tensor cores at the reference's precision (fp16 hi/lo split operands, three MMAs
                     per product, fp32 accumulation in TMEM): ~1e-6 -- the default of the planner;
`precision="bf16"`: layer 2 on the tensor cores with bf16 operands: 3x fewer MMAs, 2e-2 per evaluation -- an explicitly
                    lower-precision extra.
"""
import ctypes
import hashlib
import os
import subprocess

import numpy as np
import torch

from . import build as pgbuild
from .lib import PygentB200Error

PRECISIONS = {"fp32": 0, "bf16": 1, "f16x3": 2}
TENSOR_CORE = ("bf16", "f16x3")
_CU = [os.path.join(pgbuild.CSRC, "pg_mlp.cu"), os.path.join(pgbuild.CSRC, "pg_train.cu")]
_SRC = _CU + [os.path.join(pgbuild.CSRC, "pg_math.cuh"), os.path.join(pgbuild.CSRC, "pg_tc.cuh"),
              os.path.join(pgbuild.INCLUDE, "pygent_b200_mlp.h")]


class Dims(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int32), ("m", ctypes.c_int32), ("n_obs", ctypes.c_int32), ("dx_dim", ctypes.c_int32),
                ("is_angle", ctypes.c_uint8 * 8)]


def library_path():
    h = hashlib.sha1()
    for p in _SRC:
        with open(p, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(pgbuild.NVCC_FLAGS).encode())
    return os.path.join(pgbuild.CACHE, "pg_mlp_%s.so" % h.hexdigest()[:12])


def build_library(force=False):
    so = library_path()
    if os.path.isfile(so) and not force:
        return so
    nvcc = pgbuild.find_nvcc()
    if nvcc is None:
        raise RuntimeError("pygent_b200: %s is missing and nvcc was not found; there is no CPU fallback" % so)
    os.makedirs(pgbuild.CACHE, exist_ok=True)
    tmp = so + ".tmp%d" % os.getpid()
    res = subprocess.run([nvcc] + pgbuild.NVCC_FLAGS + _CU + ["-o", tmp], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed for pg_mlp.cu / pg_train.cu:\n" + res.stderr[-4000:])
    os.replace(tmp, so)
    return so


_vp, _i32, _i64, _f64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
_fp = ctypes.POINTER(ctypes.c_float)
_dims_p = ctypes.POINTER(Dims)
SIGNATURES = {
    "pg_mlp_pack": [_dims_p] + [_fp] * 12 + [_vp],
    "pg_mlp_ode": [_i32, _dims_p, _vp, _i64, _f64, _vp, _vp, _vp, _vp],
    "pg_mlp_rollout": [_i32, _dims_p, _vp, _i64, _i32, _f64, _vp, _vp, _vp, _vp, _vp, _vp, _f64,
                       ctypes.POINTER(ctypes.c_double), _vp, _vp, _vp],
}
SIGNATURES["pg_mlp_forward"] = [_i32, _dims_p, _vp, _i64, _i32, _f64, _i32, ctypes.POINTER(ctypes.c_double), _vp, _vp, _vp, _vp,
                              _vp, _vp, ctypes.POINTER(ctypes.c_double), _vp, _vp, _vp, _vp, _vp]
SIGNATURES["pg_mlp_forward_balanced"] = SIGNATURES["pg_mlp_forward"][:-1] + [_vp, _i64, _vp]
SIGNATURES["pg_mlp_linearize"] = [_i32, _i32, _dims_p, _vp, _i64, _i32, _f64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
_f32 = ctypes.c_float
SIGNATURES["pg_mlp_train_init"] = [_dims_p, _vp, _vp, _i64, _i32, _vp]
SIGNATURES["pg_mlp_train_grad"] = [_dims_p, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _vp, _i64, _i32, _vp, _vp]
SIGNATURES["pg_mlp_train_update"] = [_dims_p, _vp, _vp, _vp, _f32, _f32, _f32, _vp, _i64, _i32, _vp]
SIGNATURES["pg_mlp_train_p2p_alloc"] = [_i64, ctypes.POINTER(ctypes.c_void_p), _vp]
SIGNATURES["pg_mlp_train_p2p_open"] = [_vp, ctypes.POINTER(ctypes.c_void_p)]
SIGNATURES["pg_mlp_train_p2p_close"] = [_vp, _i32]
SIGNATURES["pg_mlp_train_update_p2p"] = [_dims_p, _vp, _vp, ctypes.POINTER(ctypes.c_void_p), _i32, _i32, _i64, _f32, _f32, _f32, _vp, _vp, _i64, _i32, _vp]
SIGNATURES["pg_mlp_train_step"] = [_dims_p, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _f32, _f32, _f32, _vp, _i64, _i32, _vp, _vp]
LIN_MODES = {"analytic": 0, "fd": 1}
_dll = None


def dll():
    global _dll
    if _dll is None:
        _dll = ctypes.CDLL(build_library())
        for name, argtypes in SIGNATURES.items():
            fn = getattr(_dll, name)
            fn.argtypes, fn.restype = argtypes, ctypes.c_int
        _dll.pg_mlp_packed_bytes.restype = ctypes.c_int64
        _dll.pg_mlp_forward_workspace_bytes.argtypes, _dll.pg_mlp_forward_workspace_bytes.restype = [_i64, _i32], ctypes.c_int64
        _dll.pg_mlp_train_param_count.argtypes, _dll.pg_mlp_train_param_count.restype = [_dims_p], ctypes.c_int64
        _dll.pg_mlp_train_workspace_bytes.argtypes, _dll.pg_mlp_train_workspace_bytes.restype = [_i32], ctypes.c_int64
        _dll.pg_mlp_train_p2p_bytes.argtypes, _dll.pg_mlp_train_p2p_bytes.restype = [_dims_p], ctypes.c_int64
    return _dll


def _check(rc, what):
    if rc:
        raise PygentB200Error("%s failed: %s" % (what, "bad argument" if rc < 0 else "CUDA error %d" % rc))


def _dev(a, device):
    """host [R, c] -> device [c, R] fp64"""
    return torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=float).T), dtype=torch.float64, device=device)


def _ptr(t):
    if t is None:
        return None
    if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float64):
        raise PygentB200Error("MLP path expects contiguous fp64 CUDA tensors")
    return ctypes.c_void_p(t.data_ptr())


class MLPDynamics:
    """Packed network + moments on the device."""

    def __init__(self, dims, packed, device):
        self.dims, self.packed, self.device = dims, packed, device
        self.n, self.m = dims.n, dims.m
        self.launches = 0
        # tensor-core rollouts on the load-balanced schedule (pg_mlp_forward_balanced; bit-identical to the plain one)
        self.balanced = os.environ.get("PG_MLP_BALANCED", "1") != "0"
        self._ws = {}

    def _workspace(self, B, n_alpha):
        """Progress words of a balanced rollout: one buffer per stream (calls on one stream are ordered)."""
        nb = int(dll().pg_mlp_forward_workspace_bytes(B, n_alpha))
        key = torch.cuda.current_stream(self.device).cuda_stream
        ws = self._ws.get(key)
        if ws is None or ws.numel() * 4 < nb:
            ws = self._ws[key] = torch.zeros((max(nb // 4, 1024),), dtype=torch.int32, device=self.device)
        return ws, ws.numel() * 4

    @classmethod
    def from_arrays(cls, n, m, is_angle, W1, b1, W2, b2, W3, b3, o_mean=None, o_var=None, u_mean=None, u_var=None,
                    dx_mean=None, dx_var=None, device=None):
        if device is None:
            if not torch.cuda.is_available():
                raise PygentB200Error("pygent_b200 needs a CUDA device (B200); there is no CPU fallback")
            device = torch.device("cuda", torch.cuda.current_device())
        d = Dims()
        d.n, d.m = n, m
        d.n_obs = int(n + np.sum(np.asarray(is_angle, dtype=bool)))
        d.dx_dim = n // 2
        for i, a in enumerate(is_angle):
            d.is_angle[i] = 1 if a else 0
        f = lambda a: None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32).ravel())
        arrs = [f(a) for a in (W1, b1, W2, b2, W3, b3, o_mean, o_var, u_mean, u_var, dx_mean, dx_var)]
        nbytes = int(dll().pg_mlp_packed_bytes())
        host = torch.empty(nbytes + 1024, dtype=torch.uint8)
        off = (-host.data_ptr()) % 1024
        ptrs = [None if a is None else a.ctypes.data_as(_fp) for a in arrs]
        _check(dll().pg_mlp_pack(ctypes.byref(d), *ptrs, ctypes.c_void_p(host.data_ptr() + off)), "pg_mlp_pack")
        raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
        doff = (-raw.data_ptr()) % 1024
        packed = raw[doff:doff + nbytes]
        packed.copy_(host[off:off + nbytes])
        obj = cls(d, packed, device)
        obj._raw = raw
        return obj

    @classmethod
    def from_module(cls, net, device=None):
        """net: `pygent_b200.nn_models.NNDynamics` (or the reference's): layer1/2/3 + moments."""
        g = lambda t: t.detach().cpu().numpy()
        assert net.dxDim == net.xDim // 2, "only sparse_dyn networks (velocity half, mbrl.py:54-57) are supported"
        return cls.from_arrays(net.xDim, net.uDim, net.xIsAngle, g(net.layer1.weight), g(net.layer1.bias),
                               g(net.layer2.weight), g(net.layer2.bias), g(net.layer3.weight), g(net.layer3.bias),
                               g(net.oMean), g(net.oVar), g(net.uMean), g(net.uVar), g(net.dxMean), g(net.dxVar), device)

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def ode_device(self, x, u, dt, precision="fp32"):
        """x [n, R], u [m, R] device fp64 -> rhs [n, R] = MBRL.ode."""
        R = x.shape[1]
        rhs = torch.empty_like(x)
        _check(dll().pg_mlp_ode(PRECISIONS[precision], ctypes.byref(self.dims), ctypes.c_void_p(self.packed.data_ptr()),
                                R, dt, _ptr(x), _ptr(u), _ptr(rhs), self._stream()), "pg_mlp_ode")
        self.launches += 1
        return rhs

    def ode(self, x, u, dt, precision="fp32"):
        """host x [R, n], u [R, m] -> rhs [R, n]"""
        rhs = self.ode_device(_dev(x, self.device), _dev(u, self.device), dt, precision)
        return rhs.cpu().numpy().T

    def rollout_device(self, x0, dt, T=None, U=None, gains=None, alpha=1.0, ulim=None, precision="fp32", want_u=False):
        """x0 [n, B]; open loop U [T, m, B] (None: u = 0) or closed loop gains = (KK, kk, xbar, ubar).
        Returns X [T+1, n, B] (and the applied controls [T, m, B] if want_u)."""
        B = x0.shape[1]
        if U is not None:
            T = U.shape[0]
        if gains is not None:
            T = gains[3].shape[0]
        X = torch.empty((T + 1, self.n, B), dtype=torch.float64, device=self.device)
        Uo = torch.empty((T, self.m, B), dtype=torch.float64, device=self.device) if want_u else None
        KK, kk, xbar, ubar = gains if gains is not None else (None, None, None, None)
        lim = None if ulim is None else (ctypes.c_double * len(ulim))(*[float(v) for v in ulim])
        if precision in TENSOR_CORE and self.balanced:
            ws, nb = self._workspace(B, 1)
            al = (ctypes.c_double * 1)(float(alpha))
            _check(dll().pg_mlp_forward_balanced(PRECISIONS[precision], ctypes.byref(self.dims), ctypes.c_void_p(self.packed.data_ptr()),
                                                 B, T, dt, 1, al, _ptr(x0), _ptr(U), _ptr(KK), _ptr(kk), _ptr(xbar), _ptr(ubar),
                                                 lim, _ptr(X), _ptr(Uo), None, None, ctypes.c_void_p(ws.data_ptr()), nb, self._stream()),
                   "pg_mlp_forward_balanced")
        else:
            _check(dll().pg_mlp_rollout(PRECISIONS[precision], ctypes.byref(self.dims), ctypes.c_void_p(self.packed.data_ptr()),
                                        B, T, dt, _ptr(x0), _ptr(U), _ptr(KK), _ptr(kk), _ptr(xbar), _ptr(ubar), alpha, lim,
                                        _ptr(X), _ptr(Uo), self._stream()), "pg_mlp_rollout")
        self.launches += 1
        return (X, Uo) if want_u else X

    def forward_device(self, x0, dt, alphas, gains, ulim=None, precision="bf16", out=None, work=None):
        """iLQR.forward_pass for every alpha in one launch: gains = (KK, kk, xbar, ubar) in kernel layout.
        Returns (X [A,T+1,n,B], U [A,T,m,B]); `work` = (index i32 [B], count i32 [1]) restricts the trajectories."""
        B = x0.shape[1]
        KK, kk, xbar, ubar = gains
        T, A = ubar.shape[0], len(alphas)
        out = out or {}
        X = out.get("X") if out.get("X") is not None else torch.empty((A, T + 1, self.n, B), dtype=torch.float64, device=self.device)
        Uo = out.get("U") if out.get("U") is not None else torch.empty((A, T, self.m, B), dtype=torch.float64, device=self.device)
        al = (ctypes.c_double * A)(*[float(v) for v in alphas])
        lim = None if ulim is None else (ctypes.c_double * len(ulim))(*[float(v) for v in ulim])
        idx, cnt = (None, None) if work is None else (ctypes.c_void_p(work[0].data_ptr()), ctypes.c_void_p(work[1].data_ptr()))
        if precision in TENSOR_CORE and self.balanced:
            ws, nb = self._workspace(B, A)
            _check(dll().pg_mlp_forward_balanced(PRECISIONS[precision], ctypes.byref(self.dims), ctypes.c_void_p(self.packed.data_ptr()),
                                                 B, T, dt, A, al, _ptr(x0), None, _ptr(KK), _ptr(kk), _ptr(xbar), _ptr(ubar), lim,
                                                 _ptr(X), _ptr(Uo), idx, cnt, ctypes.c_void_p(ws.data_ptr()), nb, self._stream()),
                   "pg_mlp_forward_balanced")
        else:
            _check(dll().pg_mlp_forward(PRECISIONS[precision], ctypes.byref(self.dims), ctypes.c_void_p(self.packed.data_ptr()), B, T,
                                        dt, A, al, _ptr(x0), None, _ptr(KK), _ptr(kk), _ptr(xbar), _ptr(ubar), lim, _ptr(X), _ptr(Uo),
                                        idx, cnt, self._stream()), "pg_mlp_forward")
        self.launches += 1
        return X, Uo

    def linearize_device(self, xbar, ubar, dt, mode="analytic", precision="bf16", out=None, work=None):
        """Continuous-time Jacobians of MBRL.ode along (xbar [T+1,n,B], ubar [T,m,B]): A [T,n,n,B], Bm [T,n,m,B].
        mode "fd": the reference's central differences (eps = 1e-3, 2(n+m) network evaluations per point);
        mode "analytic": chain rule on the tensor cores (precision "bf16" only)."""
        T, B = ubar.shape[0], ubar.shape[2]
        out = out or {}
        A = out.get("A") if out.get("A") is not None else torch.empty((T, self.n, self.n, B), dtype=torch.float64, device=self.device)
        Bm = out.get("B") if out.get("B") is not None else torch.empty((T, self.n, self.m, B), dtype=torch.float64, device=self.device)
        idx, cnt = (None, None) if work is None else (ctypes.c_void_p(work[0].data_ptr()), ctypes.c_void_p(work[1].data_ptr()))
        _check(dll().pg_mlp_linearize(PRECISIONS[precision], LIN_MODES[mode], ctypes.byref(self.dims),
                                      ctypes.c_void_p(self.packed.data_ptr()), B, T, dt, _ptr(xbar), _ptr(ubar), _ptr(A), _ptr(Bm),
                                      idx, cnt, self._stream()), "pg_mlp_linearize")
        self.launches += 1
        return A, Bm


def make_dims(n, m, is_angle):
    d = Dims()
    d.n, d.m = n, m
    d.n_obs = int(n + np.sum(np.asarray(is_angle, dtype=bool)))
    d.dx_dim = n // 2
    for i, a in enumerate(is_angle):
        d.is_angle[i] = 1 if a else 0
    return d


class TrainKernels:
    """Device state and launches of the dynamics training step (include/pygent_b200_mlp.h, "dynamics training"; reference
    pygent/algorithms/mbrl.py:450-509, Adagrad :71-73): flat fp32 `params` / `state_sum` / `bucket` buffers in torch's
    parameter order, the operand-image workspace of the tensor-core GEMMs, the data set and its moments on the device.
    There is no host path: construction fails without a CUDA device."""

    def __init__(self, dims, params, max_rows=512, device=None):
        if device is None or torch.device(device).type != "cuda" or not torch.cuda.is_available():
            raise PygentB200Error("the dynamics training step runs on the B200 only; there is no CPU fallback")
        self.device = torch.device(device)
        self.dims, self.max_rows = dims, int(max_rows)
        self.count = int(dll().pg_mlp_train_param_count(ctypes.byref(dims)))
        if self.count <= 0 or params.numel() != self.count:
            raise PygentB200Error("flat parameter buffer: expected %d floats, got %d" % (self.count, params.numel()))
        if not (params.is_cuda and params.is_contiguous() and params.dtype == torch.float32):
            raise PygentB200Error("params must be a contiguous fp32 CUDA tensor")
        self.params = params
        self.state_sum = torch.zeros_like(params)
        self.bucket = torch.zeros(self.count + 1, dtype=torch.float32, device=self.device)  # gradient + loss
        self.ws_bytes = int(dll().pg_mlp_train_workspace_bytes(self.max_rows))
        self._raw = torch.empty(self.ws_bytes + 1024, dtype=torch.uint8, device=self.device)
        self._ws_ptr = self._raw.data_ptr() + (-self._raw.data_ptr()) % 1024
        self.moments = torch.zeros(48, dtype=torch.float32, device=self.device)
        self.data = None
        self.launches = 0
        self.repack()

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def repack(self):
        """(Re)build the fp16 hi/lo operand images of W2 from `params` -- after construction and after any outside change."""
        _check(dll().pg_mlp_train_init(ctypes.byref(self.dims), ctypes.c_void_p(self.params.data_ptr()), ctypes.c_void_p(self._ws_ptr),
                                       self.ws_bytes, self.max_rows, self._stream()), "pg_mlp_train_init")
        self.launches += 1

    def set_moments(self, o_mean, o_var, u_mean, u_var, dx_mean, dx_var):
        m = np.zeros(48, dtype=np.float32)
        m[16:32] = 1.0
        m[34:36] = 1.0
        m[40:44] = 1.0
        for off, v in ((0, o_mean), (16, o_var), (32, u_mean), (34, u_var), (36, dx_mean), (40, dx_var)):
            v = np.asarray(v, dtype=np.float32).ravel()
            m[off:off + len(v)] = v
        self.moments.copy_(torch.from_numpy(m))

    def set_data(self, col):
        """col: {"x_", "x", "o_", "u"} -> float32 [N, dim] arrays (DataSet.columns()); one upload per training call."""
        self.data = {k: torch.as_tensor(np.ascontiguousarray(col[k], dtype=np.float32), device=self.device) for k in ("x_", "x", "o_", "u")}

    def grad(self, idx, n_total, bucket=None):
        """idx: int32 CUDA tensor of data-set rows (this rank's slice of the batch).  Fills `bucket` (default: self.bucket)."""
        bucket = self.bucket if bucket is None else bucket
        R = int(idx.numel())
        if R > self.max_rows:
            raise PygentB200Error("batch of %d rows exceeds max_rows=%d" % (R, self.max_rows))
        d = self.data
        _check(dll().pg_mlp_train_grad(ctypes.byref(self.dims), ctypes.c_void_p(self.params.data_ptr()), ctypes.c_void_p(self.moments.data_ptr()),
                                       ctypes.c_void_p(d["x_"].data_ptr()), ctypes.c_void_p(d["x"].data_ptr()),
                                       ctypes.c_void_p(d["o_"].data_ptr()), ctypes.c_void_p(d["u"].data_ptr()),
                                       ctypes.c_void_p(idx.data_ptr()) if R else None, R, int(n_total), ctypes.c_void_p(self._ws_ptr),
                                       self.ws_bytes, self.max_rows, ctypes.c_void_p(bucket.data_ptr()), self._stream()),
               "pg_mlp_train_grad")
        self.launches += 5

    def update(self, lr, weight_decay, eps=1e-10):
        _check(dll().pg_mlp_train_update(ctypes.byref(self.dims), ctypes.c_void_p(self.params.data_ptr()), ctypes.c_void_p(self.state_sum.data_ptr()),
                                         ctypes.c_void_p(self.bucket.data_ptr()), lr, weight_decay, eps, ctypes.c_void_p(self._ws_ptr),
                                         self.ws_bytes, self.max_rows, self._stream()), "pg_mlp_train_update")
        self.launches += 1

    # -- data parallel over NVLink peer memory (pg_mlp_train_update_p2p) ---------------------------------------------------------
    def enable_p2p(self, group=None):
        """Allocate this rank's exchange block, swap CUDA IPC handles with the other ranks of `group` (one all_gather of 64
        bytes) and map their blocks.  Afterwards `grad` writes into the block's parity slot and `update_p2p` replaces the
        NCCL all-reduce + `update` pair.  Raises PygentB200Error if peer mapping is not possible on this node."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        nbytes = int(dll().pg_mlp_train_p2p_bytes(ctypes.byref(self.dims)))
        own, handle = ctypes.c_void_p(), (ctypes.c_uint8 * 64)()
        _check(dll().pg_mlp_train_p2p_alloc(nbytes, ctypes.byref(own), handle), "pg_mlp_train_p2p_alloc")
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.device)
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine, group=group)
        blocks = (ctypes.c_void_p * world)()
        for r in range(world):
            if r == rank:
                blocks[r] = own.value
            else:
                h = (ctypes.c_uint8 * 64)(*every[r].cpu().tolist())
                p = ctypes.c_void_p()
                _check(dll().pg_mlp_train_p2p_open(h, ctypes.byref(p)), "pg_mlp_train_p2p_open (rank %d's block)" % r)
                blocks[r] = p.value
        self._p2p = {"world": world, "rank": rank, "blocks": blocks, "own": own.value, "npad": (self.count + 1 + 255) // 256 * 256, "step": 0,
                     "nbytes": nbytes}

        class _Mem(object):  # the own block as a torch tensor (not owned by torch: freed in close_p2p)
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        self._p2p["view"] = torch.as_tensor(_Mem(own.value, nbytes // 4), device=self.device)
        self.loss_dev = torch.zeros(1, dtype=torch.float32, device=self.device)
        dist.barrier(group=group)  # nobody publishes into a block that is not mapped yet
        return self

    def p2p_bucket(self):
        """The slot of the exchange block that this step's gradient goes to (a view of count + 1 floats)."""
        p = self._p2p
        off = (p["step"] & 1) * p["npad"]
        return p["view"][off:off + self.count + 1]

    def update_p2p(self, lr, weight_decay, eps=1e-10):
        p = self._p2p
        _check(dll().pg_mlp_train_update_p2p(ctypes.byref(self.dims), ctypes.c_void_p(self.params.data_ptr()), ctypes.c_void_p(self.state_sum.data_ptr()),
                                             p["blocks"], p["world"], p["rank"], p["step"], lr, weight_decay, eps,
                                             ctypes.c_void_p(self.loss_dev.data_ptr()), ctypes.c_void_p(self._ws_ptr), self.ws_bytes, self.max_rows,
                                             self._stream()), "pg_mlp_train_update_p2p")
        p["step"] += 1
        self.launches += 2
        return self.loss_dev

    def p2p_status(self):
        """Non-zero: this rank gave up waiting for a peer's gradient (synchronises)."""
        p = self._p2p
        return int(p["view"][2 * p["npad"] + 64:2 * p["npad"] + 65].view(torch.int32).item())

    def step(self, idx, lr, weight_decay, eps=1e-10):
        """pg_mlp_train_step: the single-GPU optimiser step in one C call (six launches)."""
        R = int(idx.numel())
        d = self.data
        _check(dll().pg_mlp_train_step(ctypes.byref(self.dims), ctypes.c_void_p(self.params.data_ptr()), ctypes.c_void_p(self.state_sum.data_ptr()),
                                       ctypes.c_void_p(self.moments.data_ptr()), ctypes.c_void_p(d["x_"].data_ptr()),
                                       ctypes.c_void_p(d["x"].data_ptr()), ctypes.c_void_p(d["o_"].data_ptr()), ctypes.c_void_p(d["u"].data_ptr()),
                                       ctypes.c_void_p(idx.data_ptr()) if R else None, R, lr, weight_decay, eps, ctypes.c_void_p(self._ws_ptr),
                                       self.ws_bytes, self.max_rows, ctypes.c_void_p(self.bucket.data_ptr()), self._stream()),
               "pg_mlp_train_step")
        self.launches += 6
