"""Regenerate `pygent_b200/modeling/expr/*.json` from the Lagrangian derivations.

    python -m pygent_b200.modeling.derive [name ...]

Each file holds f(x,u), A = df/dx, B = df/du as sympy-parsable strings in the symbols
x1..xn, u (17 significant digits).  They play the role of the reference's pickled models
(`pygent/modeling_scripts/c_files/*.p`), in a format that survives sympy upgrades.
"""
import json
import os
import sys
import time

import sympy as sp

from . import lagrange

EXPR_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "expr")

VARIANTS = [(name, lin) for name in lagrange.DERIVATIONS for lin in (True, False)
            if not (name == "cart_pole_triple" and not lin)]


def stem(name, linearized):
    return name + ("_lin" if linearized else "")


def derive_one(name, linearized):
    f, xs, u = lagrange.DERIVATIONS[name](linearized)
    A = f.jacobian(list(xs))
    B = f.diff(u)
    pr = lambda e: sp.sstr(e, full_prec=True)
    return {
        "name": stem(name, linearized),
        "n": len(xs), "m": 1,
        "f": [pr(e) for e in f],
        "A": [[pr(A[i, j]) for j in range(len(xs))] for i in range(len(xs))],
        "B": [[pr(B[i, 0])] for i in range(len(xs))],
    }


def main(argv):
    os.makedirs(EXPR_DIR, exist_ok=True)
    todo = [(n, l) for n, l in VARIANTS if not argv or stem(n, l) in argv or n in argv]
    for name, lin in todo:
        t0 = time.time()
        d = derive_one(name, lin)
        with open(os.path.join(EXPR_DIR, d["name"] + ".json"), "w") as fh:
            json.dump(d, fh, indent=1)
        print("%-36s n=%d  %.1fs" % (d["name"], d["n"], time.time() - t0))


if __name__ == "__main__":
    main(sys.argv[1:])
