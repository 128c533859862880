"""Recipe for oracle/_ref: the UNMODIFIED reference (mpritzkoleit/pygent 0.15, pure Python + its sympy model pickles),
installed from /root/reference with pip, offline, into the git-ignored directory oracle/_ref/ so that it travels to the GPU
box with the rest of the tree and `bench.py`'s CPU legs can time the reference itself there.   python -m oracle.install_ref

Nothing of the reference enters the repository's history (oracle/_ref/ is in .gitignore); /root/reference is read-only, so the
build happens in a throw-away copy.  TEST / BASELINE TOOLING ONLY: the product never imports this.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("PYGENT_REFERENCE", "/root/reference")


def installed():
    return os.path.isfile(os.path.join(TARGET, "pygent", "algorithms", "ilqr.py"))


def install(force=False):
    """Returns True if oracle/_ref holds the reference afterwards."""
    if installed() and not force:
        return True
    if not os.path.isdir(os.path.join(SOURCE, "pygent")):
        return False
    tmp = tempfile.mkdtemp(prefix="pygent_ref_")
    try:
        src = os.path.join(tmp, "src")
        shutil.copytree(SOURCE, src, ignore=shutil.ignore_patterns(".git", "examples"))
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--no-compile", "--quiet",
               "--target", TARGET, src]
        res = subprocess.run(cmd, capture_output=True, text=True, cwd=tmp)
        if res.returncode != 0:
            raise RuntimeError("pip install of the reference failed:\n" + res.stderr[-2000:])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return installed()


if __name__ == "__main__":
    print("oracle/_ref:", "installed" if install(force="--force" in sys.argv) else "reference tree not available")
