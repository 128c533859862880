"""CPU-only checks: C-ABI export list, code generation, tracing, host-side API logic."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import sympy as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    from pygent_b200 import build as pgbuild
    specs = ge.standard_specs()
    paths = [pgbuild.library_path(s) for s in specs]
    if not all(os.path.isfile(p) for p in paths):
        ge.build(verbose=False)
    return specs, paths


def test_header_declares_what_the_library_exports(built):
    """Every function include/pygent_b200.h declares is exported (extern "C") by every problem library,
    and the ctypes signatures cover exactly that list.  No compute call is made (no GPU here)."""
    from pygent_b200 import lib
    hdr = open(os.path.join(ROOT, "include", "pygent_b200.h")).read()
    declared = set(re.findall(r"^int (pg_\w+)\(", hdr, flags=re.M))
    assert declared == set(lib.SIGNATURES), declared ^ set(lib.SIGNATURES)
    declared |= set(re.findall(r"^int64_t (pg_\w+)\(", hdr, flags=re.M))
    assert "pg_rollout_workspace_bytes" in declared
    specs, paths = built
    for p in paths:
        dll = ctypes.CDLL(p)
        for name in declared:
            assert hasattr(dll, name), (p, name)
    L = lib.ProblemLibrary(paths[0])
    assert (L.n, L.m, L.n_obs, L.has_ab) == (4, 1, 5, True) and L.name == "cart_pole_lin"
    assert L.info.abi_version == 7


def test_mlp_header_declares_what_the_library_exports():
    """include/pygent_b200_mlp.h vs the learned-dynamics library and its ctypes signatures (no compute call)."""
    from pygent_b200 import mlp
    hdr = open(os.path.join(ROOT, "include", "pygent_b200_mlp.h")).read()
    declared = set(re.findall(r"^int (pg_\w+)\(", hdr, flags=re.M))
    assert declared == set(mlp.SIGNATURES), declared ^ set(mlp.SIGNATURES)
    dll = ctypes.CDLL(mlp.build_library())
    for name in declared | {"pg_mlp_packed_bytes"}:
        assert hasattr(dll, name), name
    assert ctypes.sizeof(mlp.Dims) == 24


def test_select_params_layout_matches_header():
    from pygent_b200 import lib
    # int32 x4, double x16, double x6 with natural alignment
    assert ctypes.sizeof(lib.SelectParams) == 16 + 8 * 16 + 8 * 6
    assert ctypes.sizeof(lib.ProblemInfo) == 8 * 4 + 64 + 32


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pygent_b200 import lib
    from pygent_b200.examples import make_env
    env = make_env("cart_pole")
    with pytest.raises(lib.PygentB200Error):
        env.step([0.0])
    L = env.library()
    with pytest.raises(lib.PygentB200Error):
        L.rollout(torch.zeros((4, 2), dtype=torch.float64), T=1, balanced=False)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pygent_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "pygent_oracle" not in src, f


def test_fp32_kernels_contain_no_fp64_arithmetic(built):
    """The float instantiation must not silently promote to double (literals are wrapped in T(...)).
    The only FP64 instruction tolerated is the single DMUL inside CUDA's sincosf() large-argument
    reduction (I2F.F64 -> DMUL -> F2F), which the benchmark ranges never reach."""
    specs, paths = built
    sass = subprocess.run(["cuobjdump", "-sass", paths[0]], capture_output=True, text=True, check=True).stdout
    cur, bad = None, {}
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
        elif cur and re.search(r"I7ProblemfL|I7ProblemfE", cur) and re.search(r"\b(DFMA|DADD)\b", line):
            bad[cur] = bad.get(cur, 0) + 1
    assert not bad, bad


def test_models_match_reference_golden(golden):
    """The product's sympy models (own Lagrangian derivation, pygent_b200/modeling) at the reference's
    golden points: f, A, B within 1e-11."""
    from pygent_b200.modeling.systems import get_system
    g = golden("models")
    for system in sorted({k.split("/")[0] for k in g.files}):
        s = get_system(system)
        f_np, A_np, B_np = s.numpy_callables()
        xs, us = g[system + "/x"], g[system + "/u"]
        f = np.array([f_np(0, x, u) for x, u in zip(xs, us)])
        assert np.max(np.abs(f - g[system + "/f"]) / (np.abs(g[system + "/f"]) + 1)) < 2e-12, system
        if system + "/A" in g.files:
            A = np.array([A_np(x, u) for x, u in zip(xs, us)])
            Bm = np.array([B_np(x, u) for x, u in zip(xs, us)])
            assert np.max(np.abs(A - g[system + "/A"]) / (np.abs(g[system + "/A"]) + 1)) < 2e-11, system
            assert np.max(np.abs(Bm - g[system + "/B"]) / (np.abs(g[system + "/B"]) + 1)) < 2e-11, system


def test_cost_dispatch_and_tracing():
    """The reference's cost-signature dispatch (environments.py:206-235) and symbolic tracing."""
    from pygent_b200 import tracing
    xs, us, xn, t = list(sp.symbols("x1:3")), list(sp.symbols("u1:2")), list(sp.symbols("xn1:3")), sp.Symbol("t")
    cases = [
        (lambda x: x[0] ** 2, xs[0] ** 2),
        (lambda x, u: x[0] + u[0], xs[0] + us[0]),
        (lambda x, t: x[0] * t, xs[0] * t),
        (lambda x, mod: mod.cos(x[1]), sp.cos(xs[1])),
        (lambda x, u, x2: x2[0] - x[0], xn[0] - xs[0]),
        (lambda x, u, t: u[0] * t, us[0] * t),
        (lambda x, u, mod: mod.exp(u[0]), sp.exp(us[0])),
        (lambda x, u, t, mod: mod.exp(-t) * u[0], sp.exp(-t) * us[0]),
        (lambda x, u, x2, mod: mod.sin(x2[1]), sp.sin(xn[1])),
        (lambda x, u, x2, t: x2[1] * t, xn[1] * t),
        (lambda x_, u_, x, t, mod: mod.sin(x[0]) + t, sp.sin(xn[0]) + t),
    ]
    for fn, want in cases:
        got = tracing.trace(tracing.bind_cost(fn), xs, us, xn, t, sp)
        assert sp.simplify(got - want) == 0
    with pytest.raises(TypeError):
        tracing.bind_cost(lambda a, b, c, d, e, f: 0)
    # numpy ufuncs inside user code
    got = tracing.trace(lambda x, u: np.cos(x[1]) ** 2 + np.sqrt(u[0] * u[0] + 1.0), xs, us)
    assert sp.simplify(got - (sp.cos(xs[1]) ** 2 + sp.sqrt(us[0] ** 2 + 1.0))) == 0


def test_codegen_pairs_sincos_and_masks():
    from pygent_b200.codegen import ProblemSpec
    from pygent_b200.modeling.systems import get_system
    s = get_system("cart_pole_lin")
    x1, x2, x3, x4 = s.xs
    spec = ProblemSpec(s, x1 ** 2 + s.us[0] ** 2, x2 ** 2, term=[(0, 1.2)], x_is_angle=[0, 1, 0, 0])
    h = spec.header("k")
    assert h.count("m.sincos(x[1]") == 2  # once in f, once in AB
    assert "k1 = pg::hoist(PG_PROBLEM_CONSTS, 1, T(1.477886191760791));" in h and "m.k1*" in h
    assert "a_nz(int i) { constexpr unsigned char m[16] = {0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 1}" in h
    assert "NOBS = 5" in h and "HAS_AB = true" in h
    assert spec.key() != ProblemSpec(s, x1 ** 2 + 2 * s.us[0] ** 2, x2 ** 2).key()
    # a cost on the next state has no iLQR expansion
    spec2 = ProblemSpec(s, spec.xn[0] ** 2)
    assert "HAS_COST_DERIVS = false" in spec2.header("k")
    with pytest.raises(ValueError):
        ProblemSpec(s, sp.Symbol("y") ** 2)


def test_environment_host_logic():
    """Shapes, reset, terminate, observe -- nothing here needs the GPU."""
    from pygent_b200.examples import make_env
    env = make_env("cart_pole")
    assert (env.xDim, env.uDim, env.oDim) == (4, 1, 5) and not env.batched
    assert np.allclose(env.x, [0, np.pi, 0, 0]) and env.history.shape == (1, 4) and env.tt == [0]
    assert env.terminate([1.3, 0, 0, 0]) and not env.terminate([1.0, 0, 3.9, 0]) and env.terminate([0, 0, -4.1, 0])
    assert np.allclose(env.observe([0.5, np.pi, 1, 2]), [0.5, -1, 0, 1, 2])
    assert float(env.uMax[0]) == 10.0
    envb = make_env("double_parallel", x0=np.zeros((5, 6)))
    assert envb.batched and envb.batch == 5 and envb.x.shape == (5, 6) and envb.history.shape == (5, 1, 6)
    assert envb.terminate(np.array([[0, 0, 0, 0, 26, 0], [0, 0, 0, 0, 0, 0]])).tolist() == [True, False]
    envb.reset(np.ones((3, 6)))
    assert envb.batch == 3 and envb.x.shape == (3, 6)
    env.reset(lambda: [0.1, 0.2, 0.3, 0.4])
    assert np.allclose(env.x, [0.1, 0.2, 0.3, 0.4])
    with pytest.raises(ValueError):
        from pygent_b200.environments import StateSpaceModel
        StateSpaceModel("cart_pole_lin", lambda x, u: 0, [0, 0], 1, 0.02)


def test_oracle_threads_agree():
    """The OpenMP batch loops of the oracle give the same bits as the serial ones (cpu_baseline uses them)."""
    import oracle
    from cases import oracle_problem
    p = oracle_problem("cart_pole")
    rng = np.random.default_rng(0)
    x0 = np.array([0, np.pi, 0, 0]) + rng.uniform(-0.1, 0.1, (16, 4))
    U = rng.uniform(-5, 5, (16, 30, 1))
    a, b = oracle.rollout(p, x0, U=U, nthreads=1), oracle.rollout(p, x0, U=U, nthreads=4)
    assert np.array_equal(a["X"], b["X"]) and np.array_equal(a["cost"], b["cost"])
    q = oracle.ilqr_params(max_iters=5)
    sa, sb = oracle.ilqr_solve(p, q, x0[:4], 30, nthreads=1), oracle.ilqr_solve(p, q, x0[:4], 30, nthreads=4)
    assert np.array_equal(sa["xx"], sb["xx"]) and np.array_equal(sa["cost"], sb["cost"])


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py's stdout contract: ONE JSON line, whatever else writes to file descriptor 1 (the reference arm runs on the
    host cores, so this is checkable without a GPU)."""
    import json
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "rk4_env_steps_per_s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


def test_host_action_stream_known_answers():
    """pygent_b200/random_actions.py: Philox2x32-10 (the action stream's generator) and Philox4x32-10 against the Random123
    known-answer vectors (kat_vectors: philox2x32 10, philox4x32 10), and the shape / range / determinism of the action stream."""
    from pygent_b200.random_actions import philox2x32_10, philox4x32_10, random_actions
    for ctr, key, want in [((0, 0), 0, (0xff1dae59, 0x6cd10df2)), ((0xffffffff, 0xffffffff), 0xffffffff, (0x2c3f628b, 0xab4fd7ad)),
                           ((0x243f6a88, 0x85a308d3), 0x13198a2e, (0xdd7ce038, 0xf62a4c12))]:
        assert tuple(int(g) for g in philox2x32_10(*ctr, key)) == want
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        got = philox4x32_10(*ctr, *key)
        assert tuple(int(g) for g in got) == want
    U = random_actions(257, 40, seed=12345, u_max=[10.0, 0.5])
    assert U.shape == (40, 2, 257) and np.all(np.abs(U[:, 0]) <= 10.0) and np.all(np.abs(U[:, 1]) <= 0.5)
    assert abs(U[:, 0].mean()) < 0.2 and abs(U[:, 0].std() - 10 / np.sqrt(3)) < 0.2
    assert np.array_equal(U[10:], random_actions(257, 30, seed=12345, u_max=[10.0, 0.5], k_offset=10))      # the stream is indexed, not stateful
    assert np.array_equal(U[:, :, 100:], random_actions(157, 40, seed=12345, u_max=[10.0, 0.5], b_offset=100))
    assert not np.array_equal(U, random_actions(257, 40, seed=12346, u_max=[10.0, 0.5]))
