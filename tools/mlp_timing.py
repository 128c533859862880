"""Where an interval of mlp_x3_kernel goes: builds pg_mlp.cu with -DPG_MLP_TIMING (cycle counters per phase, see the
macro's comment in csrc/pg_mlp.cu) into pygent_b200/_problems/pg_mlp_timing.so and runs a 45,056-row line search on it.
    python tools/mlp_timing.py build     # here (nvcc, no GPU)
    python tools/mlp_timing.py           # on the GPU box"""
import ctypes
import os
import subprocess
import sys

import numpy as np

sys.path[:0] = [".", "tests"]
from pygent_b200 import build as pgbuild
from pygent_b200 import mlp

SO = os.path.join(pgbuild.CACHE, "pg_mlp_timing.so")

if len(sys.argv) > 1 and sys.argv[1] == "build":
    cmd = [pgbuild.find_nvcc()] + pgbuild.NVCC_FLAGS + ["-DPG_MLP_TIMING"] + sys.argv[2:] + mlp._CU + ["-o", SO]
    subprocess.run(cmd, check=True)
    print("built", SO)
    sys.exit(0)

import torch
from oracle import mlp_oracle as mo

mlp.library_path = lambda: SO
w, mom = mo.make_weights(6, 2), mo.make_moments(5, 1, 2)
dyn = mlp.MLPDynamics.from_arrays(4, 1, [False, True, False, False], w["W1"], w["b1"], w["W2"], w["b2"], w["W3"], w["b3"], **mom)
rng = np.random.default_rng(0)
R, T = (int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 45056), 100
x0 = torch.as_tensor(rng.uniform(-1, 1, (4, R)), device="cuda")
U = torch.as_tensor(rng.uniform(-6, 6, (T, 1, R)), device="cuda")
dll = ctypes.CDLL(SO)
buf = (ctypes.c_uint64 * (1024 * 16))()
for _ in range(2):
    dyn.rollout_device(x0, 0.02, U=U, precision="f16x3")
torch.cuda.synchronize()
dll.pg_mlp_timing_read(buf, 1)
dyn.rollout_device(x0, 0.02, U=U, precision="f16x3")
torch.cuda.synchronize()
dll.pg_mlp_timing_read(buf, 0)
a = np.array(buf, dtype=np.float64).reshape(1024, 16)
a = a[a[:, 15] > 0]
n = a[:, 15].sum()
names = {0: "owner phase + inputs", 1: "8 layer-1 blocks", 5: "waits for a free A buffer", 2: "wait for the last MMAs", 3: "epilogue",
         4: "partial sums + update", 8: "issuer: wait A block", 9: "issuer: wait own weight stage", 10: "issuer: issuing", 11: "issuer: wait the peer's stage (relay)"}
print("rows %d: CTAs %d, intervals per CTA %.1f; cycles per interval (mean over CTAs):" % (R, len(a), a[:, 15].mean()))
tot = 0.0
for k in (0, 1, 5, 2, 3, 4):
    v = a[:, k].sum() / n
    tot += v
    print("  %-42s %8.0f" % (names[k], v))
print("  %-42s %8.0f  (%.2f us at 1.9 GHz)" % ("sum", tot, tot / 1900))
lead = a[a[:, 10] > 0]
for k in (8, 9, 11, 10):
    print("  %-42s %8.0f  (two issuing threads: mean)" % (names[k], lead[:, k].sum() / lead[:, 15].sum() / 2))
