"""does a second host thread + context on the SAME GPU close the gaps between small kernels? (development aid)"""
import importlib, sys, time, os, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("urlearning-cpp_b200")
D = importlib.import_module("urlearning-cpp_b200.distributed")
codes, card, edges, _ = pkg.datagen.discrete_bn(p=60, n=1_000_000, seed=4)
nbs = [pkg.two_hop_neighbors(edges, 60, v) for v in range(60)]
costs = [D.family_cost(card, v, nbs[v], 11) for v in range(60)]
for T in (1, 2, 3):
    engs = [pkg.Engine(0) for _ in range(T)]
    for e in engs:
        e.set_discrete(codes, card)
    owner = D.assign_lpt(costs, T)
    def work(t):
        for v in range(60):
            if owner[v] == t:
                engs[t].score_variable(v, nbs[v], 11, pkg.BIC, flags=pkg.PRUNE_DOMINATED).free()
        engs[t].synchronize()
    for rep in range(4):
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(t,)) for t in range(T)]
        [x.start() for x in th]; [x.join() for x in th]
        print(f"T={T} pass {rep}: {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)
    for e in engs:
        e.close()
