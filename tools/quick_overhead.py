import importlib, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("urlearning-cpp_b200")
codes, card, edges, _ = pkg.datagen.discrete_bn(p=60, n=1_000_000, seed=4)
eng = pkg.Engine(0)
eng.set_discrete(codes, card)
K = 11
vs = list(range(20, 40))
def run(tag, flags=0, fetch=False):
    t0 = time.time()
    for v in vs:
        res = eng.score_variable(v, pkg.two_hop_neighbors(edges, 60, v), K, pkg.BIC, flags=flags)
        if fetch: res.fetch()
        res.free()
    eng.synchronize()
    print(tag, f"{(time.time()-t0)*1e3:.1f} ms")
run("warm1"); run("warm2")
run("timing off")
eng.enable_timing(True); run("timing on"); print(eng.stats()); eng.enable_timing(False)
run("timing off again")
run("prune", flags=pkg.PRUNE_DOMINATED)
run("prune+fetch", flags=pkg.PRUNE_DOMINATED, fetch=True)
