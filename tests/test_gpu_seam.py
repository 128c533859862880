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
