"""Compile one `ProblemSpec` into its shared library with nvcc (sm_100a only).

The analogue of `sympy_to_c.convert_to_c(..., cfilepath=..., use_exisiting_so=False)` in the
reference (`pygent/modeling_scripts/cart_pole.py:159-170`): generate code, compile, load.  Libraries
are cached IN-TREE under `pygent_b200/_problems/` keyed by a hash of the expressions and of the
kernel sources, so that a library built in the CPU container travels to the GPU box.
There is no fallback: if the library is missing and nvcc is not available, this raises.
"""
import hashlib
import os
import shutil
import subprocess

from . import codegen

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
CACHE = os.path.join(HERE, "_problems")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
              "-shared", "-Xcompiler", "-fPIC"]

_SOURCES = [os.path.join(CSRC, f) for f in ("pg_abi.cu", "pg_kernels.cuh", "pg_math.cuh")] + \
           [os.path.join(INCLUDE, "pygent_b200.h")]


def _source_hash():
    h = hashlib.sha1()
    for p in _SOURCES:
        with open(p, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:12]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    return None


def problem_key(spec):
    return spec.key(extra=_source_hash())


def library_path(spec):
    return os.path.join(CACHE, "pg_%s_%s.so" % (spec.system.name, problem_key(spec)))


def build_problem(spec, force=False, verbose=False, extra_flags=()):
    """Returns the path of the compiled library for `spec`, building it if necessary."""
    key = problem_key(spec)
    so = library_path(spec)
    if os.path.isfile(so) and not force:
        return so
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError(
            "pygent_b200: compiled problem library %s is missing and nvcc was not found; "
            "there is no CPU fallback (run __graft_entry__.build() where nvcc exists)" % so)
    os.makedirs(CACHE, exist_ok=True)
    hdr = os.path.join(CACHE, "pg_%s_%s.cuh" % (spec.system.name, key))
    hdr_tmp = hdr + ".tmp%d" % os.getpid()  # several ranks may build the same problem at once: never expose a half-written header
    with open(hdr_tmp, "w") as fh:
        fh.write(spec.header(key))
    os.replace(hdr_tmp, hdr)
    tmp = so + ".tmp%d" % os.getpid()
    cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + ["-DPG_PROBLEM_HEADER=\"%s\"" % hdr,
                                                      os.path.join(CSRC, "pg_abi.cu"), "-o", tmp]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed for problem %s:\n%s\n%s" % (spec.system.name, " ".join(cmd), res.stderr[-4000:]))
    if verbose:
        print(res.stderr)
    os.replace(tmp, so)
    return so
