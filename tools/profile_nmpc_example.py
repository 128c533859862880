"""cProfile of one receding-horizon example (host side): python tools/profile_nmpc_example.py triple 10"""
import cProfile, pstats, sys, time
sys.path.insert(0, ".")
from pygent_b200.examples import make_nmpc_example

name, steps = sys.argv[1], int(sys.argv[2])
t = time.time()
alg = make_nmpc_example(name, maxIters=500)
print("construction + initial solve %.1f s" % (time.time() - t), flush=True)
pr = cProfile.Profile()
pr.enable()
alg.run(steps=steps, printing=False)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
