"""On-disk controller interchange with the UNMODIFIED reference: a cart-pole iLQR controller as the reference's
`iLQR.save()` writes it (ilqr.py:646-656: K_.npy, k.npy, uu.npy, xx.npy, alpha.npy, time_info.p) is committed under
tests/golden/ilqr_disk/data/, together with what the reference's `run_disk(x0)` (ilqr.py:353-396) makes of it from
another initial state at the controller's own step size.  (Its re-sampling branch for a different step size cannot
be pinned: `np.arange(t0, tf, dt)` there yields one sample more than there are gains whenever the accumulated final time
exceeds T*dt by rounding -- as it does here -- and `interp1d` raises.)
    python -m oracle.make_golden_disk"""
import os
import shutil
import tempfile

import numpy as np

from . import ref_harness as rh
from .make_golden import GOLDEN
from .make_golden_nmpc import c_k, c_N


def main():
    pg = rh.load_reference()
    import pygent.algorithms.ilqr as ref_ilqr
    ref_ilqr.copyfile = lambda *a, **k: None
    x0, x0b = [0.3, 0.5, 0.0, 0.0], [0.25, 0.55, 0.1, -0.1]
    out = {"x0": np.array(x0), "x0b": np.array(x0b)}
    dst = os.path.join(GOLDEN, "ilqr_disk", "data")
    os.makedirs(dst, exist_ok=True)
    with rh.fixed_step(4), rh.quiet(), tempfile.TemporaryDirectory() as tmp:
        path = tmp + "/"
        env = pg.environments.CartPole(c_k, x0, 0.02)
        alg = ref_ilqr.iLQR(env, 1.0, 0.02, fcost=c_N, constrained=False, path=path, printing=False, final_backpass=False)
        alg.run_optim()   # saving=True: writes the controller at the end
        for f in ("K_.npy", "k.npy", "uu.npy", "xx.npy", "alpha.npy", "time_info.p"):
            shutil.copyfile(path + "data/" + f, os.path.join(dst, f))
        out.update(cost=float(alg.cost))
        for tag, dt in (("same", 0.02),):
            envb = pg.environments.CartPole(c_k, x0b, dt)
            algb = ref_ilqr.iLQR(envb, 1.0, dt, fcost=c_N, constrained=False, path=path, printing=False, final_backpass=False,
                                 init=False, saving=False)
            algb.run_disk(x0b)
            out.update({tag + "_xx": np.array(algb.xx, dtype=float), tag + "_uu": np.array(algb.uu, dtype=float),
                        tag + "_cost": float(algb.cost), tag + "_KK": np.array(algb.KK, dtype=float)})
    np.savez_compressed(os.path.join(GOLDEN, "ilqr_disk.npz"), **out)
    print("ilqr_disk: cost %.6f; run_disk cost %.6f (%d steps)" % (out["cost"], out["same_cost"], len(out["same_uu"])))


if __name__ == "__main__":
    main()
