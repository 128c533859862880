"""The reference's native-codegen seam bound to the B200 implementation: the UNMODIFIED reference (oracle/_ref, the
pip-installed copy that travels with the tree) imports `sympy_to_c` from pygent_b200/shims, builds `CartPole` and `iLQR`
through `convert_to_c` (modeling_scripts/cart_pole.py:159-170, algorithms/ilqr.py:558-571), and its `ode / A / B` and cost
expansion are checked against the known answers of SURVEY.md 8(c) (measured with the reference's own lambdify path)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_unmodified_reference_binds_convert_to_c(tmp_path):
    from oracle import ref_harness as rh
    if rh.reference_root() is None:
        pytest.skip("oracle/_ref not installed (python -m oracle.install_ref where /root/reference exists)")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "seam_driver.py"), str(tmp_path)], capture_output=True, text=True,
                         timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    out = json.loads(res.stdout.strip().splitlines()[-1])
    assert out["log_has_c_function"] and not out["log_has_lambdify"], out  # the reference took its "C function" branch
    assert out["log_cost_c"] and out["ilqr_error"] is None, out
    # SURVEY 8(c) KATs at x = [0.1, 0.7, -0.3, 1.1], u = [2.5]
    assert np.allclose(out["f"], [-0.3, 1.1, 2.5, 12.096269276790501], rtol=1e-13, atol=0)
    A = np.array(out["A"])
    assert abs(A[3, 1] - 8.708529569078923) < 1e-12 and abs(A[3, 3] + 0.0631945052627076) < 1e-15
    assert A[0, 2] == 1 and A[1, 3] == 1 and np.count_nonzero(A) == 4
    assert np.allclose(out["B"], [0, 0, 1, 1.1303497074638655], rtol=1e-13, atol=0)
    # cost expansion of c_k * dt (ilqr.py:539-555): quadratic cost of examples/nmpc/cart_pole.py
    x, dt = np.array([0.1, 0.7, -0.3, 1.1]), 0.02
    assert np.allclose(out["cx"], np.array([20, 20, .01, .01]) * x * dt, rtol=1e-13)
    assert np.allclose(out["cu"], [0.02 * 2.5 * dt], rtol=1e-13)
    assert np.allclose(out["Cxx"], np.diag([20, 20, .01, .01]) * dt, rtol=1e-13)
    assert np.allclose(out["Cuu"], [[0.02 * dt]], rtol=1e-13)
    assert np.allclose(out["cfx"], np.array([200, 200, 20, 20]) * x, rtol=1e-13)
    assert np.allclose(out["Cfxx"], np.diag([200, 200, 20, 20]), rtol=1e-13)
    assert out["batched_shape"] == [4, 1, 1000] and out["batched_err"] < 1e-13


USER_SCRIPT = '''
# a script written against the REFERENCE's package names and call sequence (the shape of its examples/ilqr/*.py and
# examples/nmpc/*.py: environment, iLQR / NMPC, run, plot, animation)
from pygent.environments import CartPole, Environment
from pygent.algorithms.ilqr import iLQR
from pygent.algorithms.nmpc import NMPC
import numpy as np
import json, sys

def c_k(x, u, t, mod):
    x1, x2, x3, x4 = x
    u1, = u
    return 12*x1**2 + 6*x2**2 + 0.02*x3**2 + 0.02*x4**2 + 0.04*(500*mod.exp(-t*8) + 1)*u1**2

def c_N(x):
    x1, x2, x3, x4 = x
    return 90*x1**2 + 110*x2**2 + 12*x3**2 + 9*x4**2

x0 = [0, np.pi, 0, 0]
t, dt = 2.0, 0.02
env = CartPole(c_k, x0, dt)
env.uMax = 4*env.uMax
assert isinstance(env, Environment)
path = sys.argv[1] + '/ilqr/'
algorithm = iLQR(env, t, dt, path=path, fcost=c_N, constrained=True, maxIters=40)
xx, uu, cost = algorithm.run_optim()
algorithm.plot()
algorithm.animation()

def c2(x, u):
    x1, x2, x3, x4 = x
    u1, = u
    return 10*x1**2 + 10*x2**2 + 0.005*x3**2 + 0.005*x4**2 + 0.01*u1**2
mpc = NMPC(CartPole(c2, [0, 0.2, 0, 0], dt), CartPole(c2, [0, 0.2, 0, 0], dt), 0.4, dt, 1.0, path=sys.argv[1] + '/nmpc/', fcost=c_N,
           maxIters=20, step_iterations=1)
mpc.run()
mpc.plot()
print(json.dumps({"cost": float(cost), "xx": np.asarray(xx).shape, "uu_max": float(np.max(np.abs(uu))), "iterations": int(algorithm.iterations),
                  "plant_steps": len(mpc.environment.history), "theta_end": float(mpc.environment.history[-1][1]),
                  "module": CartPole.__module__}))
'''


def test_pygent_alias_runs_a_reference_style_script(tmp_path):
    """pygent_b200/shims/pygent: with the shims directory on PYTHONPATH a script written against the reference's package
    (`from pygent.environments import ...`, `from pygent.algorithms.ilqr import iLQR`, `.plot()`, `.animation()`) runs
    unchanged on the B200 kernels, and yields what the same calls on pygent_b200 yield."""
    script = tmp_path / "user_script.py"
    script.write_text(USER_SCRIPT)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, os.path.join(ROOT, "pygent_b200", "shims"), os.environ.get("PYTHONPATH", "")]))
    res = subprocess.run([sys.executable, str(script), str(tmp_path)], capture_output=True, text=True, timeout=900, cwd=str(tmp_path), env=env)
    assert res.returncode == 0, res.stderr[-3000:]
    out = json.loads(res.stdout.strip().splitlines()[-1])
    assert out["module"] == "pygent_b200.environments" and out["xx"] == [101, 4] and out["uu_max"] <= 40.0 + 1e-9
    assert out["plant_steps"] == 21 and abs(out["theta_end"]) < 0.2
    assert os.path.isdir(str(tmp_path / "ilqr" / "data"))
    # the same solve through pygent_b200's own names
    from pygent_b200.algorithms.ilqr import iLQR
    from pygent_b200.environments import CartPole
    ns = {}
    exec(USER_SCRIPT.split("x0 = [0, np.pi, 0, 0]")[0].split("import json, sys")[1], {"np": np}, ns)
    e = CartPole(ns["c_k"], [0, np.pi, 0, 0], 0.02)
    e.uMax = 4 * e.uMax
    alg = iLQR(e, 2.0, 0.02, path=str(tmp_path) + "/direct/", fcost=ns["c_N"], constrained=True, maxIters=40)
    _, _, cost = alg.run_optim()
    assert float(cost) == out["cost"] and int(alg.iterations) == out["iterations"]
