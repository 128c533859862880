import sys, numpy as np, torch, time
sys.path.insert(0, '/root/repo')
from pygent_b200.examples import make_ilqr, CASES
rng = np.random.default_rng(0)
x0 = np.array(CASES["triple"][4]) + rng.uniform(-0.05, 0.05, (256, 8))
alg = make_ilqr("triple", 100, x0=x0, maxIters=30)
c0 = alg.cost.copy()
t=time.time(); alg.run_optim(); torch.cuda.synchronize(); print("triple n=8: 30 iterations x 256 trajectories in %.1f ms; cost %.3f -> %.3f; finite %s" % ((time.time()-t)*1e3, c0.mean(), alg.cost.mean(), np.isfinite(alg.cost).all()))
